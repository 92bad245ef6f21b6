"""Distribution of Newton iterations (last sub-step) and line-search evaluations (whole env step) over a batch."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from gym_kmanip_b200.batch_sim import BatchSim
env = sys.argv[1] if len(sys.argv) > 1 else "KManipSoloArmQPos"
n = 4096
dtype = sys.argv[2] if len(sys.argv) > 2 else "float32"
sim = BatchSim(env, n, dtype=dtype, seed=1)
sim.configure(32, 0)
sim.reset()
gen = torch.Generator(device="cuda").manual_seed(0)
for t in range(64):
    act = torch.rand(n, sim.act_dim, device="cuda", generator=gen) * 2 - 1
    sim.step(act, contacts=True)
    if t in (2, 10, 20, 30, 45, 60):
        it, ls = sim.solver_stats()
        it, ls = it.cpu().float(), ls.cpu().float()
        ncon = sim.ncon.cpu()
        fl = sim.con_flags.cpu()
        print(f"step {t}: niter(last sub-step) mean {it.mean():.2f} max {int(it.max())} hist {torch.bincount(it.long(), minlength=8)[:12].tolist()} | "
              f"ls evals/env-step mean {ls.mean():.1f} p50 {ls.median():.0f} p99 {ls.quantile(0.99):.0f} max {int(ls.max())} | ncon hist {torch.bincount(ncon.long(), minlength=9).tolist()} coupled {(fl & 6).ne(0).float().mean():.4f}")
        w = ls.view(-1, 32)
        print(f"         per-warp(32 envs) max/mean of ls evals: {(w.max(dim=1).values / w.mean(dim=1)).mean():.2f}; niter max/mean per warp {(it.view(-1,32).max(dim=1).values / it.view(-1,32).mean(dim=1).clamp(min=0.1)).mean():.2f}")
