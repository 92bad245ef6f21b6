import sys, os
sys.path.insert(0, os.getcwd())
import torch
from gym_kmanip_b200.batch_sim import BatchSim
n=4096
sim = BatchSim("KManipSoloArmQPos", n, dtype="float32", seed=1)
sim.reset()
gen = torch.Generator(device="cuda").manual_seed(0)
prev=None
for t in range(40):
    act = torch.rand(n, sim.act_dim, device="cuda", generator=gen) * 2 - 1
    sim.step(act, contacts=False)
    it, ls = sim.solver_stats()
    ls = ls.float()
    if prev is not None and t % 6 == 0:
        c = torch.corrcoef(torch.stack([prev, ls]))[0,1].item()
        # lane efficiency if warps of 28 are formed (a) in env order (b) sorted by prev
        def eff(order):
            x = ls[order][: (n//28)*28].view(-1,28)
            return (x.mean(dim=1) / x.max(dim=1).values).mean().item(), x.max(dim=1).values.max().item(), x.max(dim=1).values.mean().item()
        e0 = eff(torch.arange(n, device=ls.device)); e1 = eff(prev.argsort()); e2 = eff(ls.argsort())
        print(f"step {t}: corr(ls[t-1], ls[t]) = {c:.2f}; mean/max per warp: env order {e0[0]:.2f} (warp max mean {e0[2]:.0f}), sorted by previous step {e1[0]:.2f} (warp max mean {e1[2]:.0f}), oracle sort {e2[0]:.2f} (warp max mean {e2[2]:.0f}); global max {e0[1]:.0f}")
    prev = ls
