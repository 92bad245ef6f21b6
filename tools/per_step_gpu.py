"""Per-step launch time and solver statistics along one episode (headline workload). usage: per_step_gpu.py [env] [n]"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from gym_kmanip_b200.batch_sim import BatchSim
env = sys.argv[1] if len(sys.argv) > 1 else "KManipSoloArmQPos"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sim = BatchSim(env, n, dtype="float32", seed=0)
sim.reset()
gen = torch.Generator(device="cuda").manual_seed(1234)
acts = torch.rand(16, n, sim.act_dim, device="cuda", generator=gen) * 2 - 1
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
prev_ls = None
for t in range(70):
    flush.fill_(t & 255)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.step(acts[t % 16], contacts=True); e1.record(); torch.cuda.synchronize()
    it, ls = sim.solver_stats()
    it = it.float(); nc = sim.ncon.float()
    dls = ls - (prev_ls if t else 0); prev_ls = ls.clone(); k = int(dls.argmax())
    w = it.view(-1, 14 if n % 14 == 0 else 32)
    print(f"step {t:2d} {e0.elapsed_time(e1):.3f} ms | niter(last sub-step) mean {it.mean():.2f} max {int(it.max())} | ncon mean {nc.mean():.2f} coupled {(sim.con_flags & 6).ne(0).float().mean():.4f} | ls evals/env-step mean {dls.float().mean():.0f} max {int(dls.max())} (env {k}: ncon {int(sim.ncon[k])} flags {int(sim.con_flags[k])}) | ncon max {int(sim.ncon.max())} n(ncon>4) {int((sim.ncon > 4).sum())}")
