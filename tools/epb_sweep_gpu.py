"""Envs-per-CTA sweep of the warp-per-env kernel (headline workload). usage: epb_sweep_gpu.py [env] [n] [epb ...]"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from gym_kmanip_b200.batch_sim import BatchSim
env = sys.argv[1] if len(sys.argv) > 1 else "KManipSoloArmQPos"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
epbs = [int(x) for x in sys.argv[3:]] or [0, 14, 10, 7, 5, 4]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for epb in epbs:
    sim = BatchSim(env, n, dtype="float32", seed=0)
    try:
        cfg = sim.configure(32, epb)
    except Exception as ex:
        print(f"epb {epb}: {ex}"); continue
    sim.reset()
    gen = torch.Generator(device="cuda").manual_seed(1234)
    acts = torch.rand(16, n, sim.act_dim, device="cuda", generator=gen) * 2 - 1
    ms = []
    for t in range(64):
        flush.fill_(t & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sim.step(acts[t % 16]); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    a, b = sum(ms[5:25]) / 20, sum(ms[40:60]) / 20
    print(f"epb {epb}: {cfg} | steps 5-25 {a:.3f} ms, steps 40-60 {b:.3f} ms", flush=True)
    sim.close()
