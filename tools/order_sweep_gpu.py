"""Cost-ordered walk A/B (KM_ORDER) over mappings. usage: order_sweep_gpu.py env n dtype lanes:epb[,lanes:epb...]"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from gym_kmanip_b200.batch_sim import BatchSim
env, n, dtype = sys.argv[1], int(sys.argv[2]), sys.argv[3]
cfgs = [tuple(int(v) for v in c.split(":")) for c in sys.argv[4].split(",")]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for lanes, epb in cfgs:
    for order in (0, 1):
        os.environ["KM_ORDER"] = str(order)
        sim = BatchSim(env, n, dtype=dtype, seed=0)
        try:
            cfg = sim.configure(lanes, epb)
        except Exception as ex:
            print(f"{lanes}:{epb}: {ex}"); sim.close(); break
        sim.reset()
        gen = torch.Generator(device="cuda").manual_seed(1234)
        acts = torch.rand(16, n, sim.act_dim, device="cuda", generator=gen) * 2 - 1
        ms = []
        for t in range(40):
            flush.fill_(t & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sim.step(acts[t % 16]); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        a = sum(ms[5:25]) / 20
        chk = float(sim.get_state()[0].double().abs().sum())
        print(f"{env} {n} {dtype} lanes {lanes} epb {cfg['envs_per_block']} grid {cfg['grid']} order {order}: {a:.3f} ms  {n / a / 1e3:.2f}e6 env-steps/s  (state checksum {chk:.6f})", flush=True)
        sim.close()
