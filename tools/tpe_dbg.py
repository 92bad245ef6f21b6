import sys, os
sys.path.insert(0, os.getcwd())
import torch
from gym_kmanip_b200.batch_sim import BatchSim
n=4096
sim = BatchSim("KManipSoloArmQPos", n, dtype="float32", seed=1)
sim.configure(1, 0)
sim.reset()
gen = torch.Generator(device="cuda").manual_seed(0)
for t in range(40):
    act = torch.rand(n, sim.act_dim, device="cuda", generator=gen) * 2 - 1
    sim.step(act, contacts=True)
    if t in (5, 20, 30, 39):
        it, kc = sim.solver_stats()
        it = it.cpu(); kc = kc.cpu().float()
        coupled = (it >> 16); nit = (it & 65535).float()
        pad = (28 - n % 28) % 28
        def warps(x, fill=0):
            x = torch.cat([x, torch.full((pad,), fill, dtype=x.dtype)]) if pad else x
            return x.view(-1, 28)
        wk = warps(kc).max(dim=1).values; wn = warps(nit); wc = warps(coupled.float())
        order = wk.argsort(descending=True)
        print(f"step {t}: per-warp kcycles mean {wk.mean():.0f} max {wk.max():.0f} min {wk.min():.0f}; newton iters/env-step mean {nit.mean():.1f} max {nit.max():.0f}; coupled sub-steps total {int(coupled.sum())}")
        for w in order[:6].tolist():
            print(f"    warp {w}: kcycles {wk[w]:.0f}  niter sum-of-lane-max? lanes niter max {wn[w].max():.0f} mean {wn[w].mean():.1f}  coupled sub-steps in warp {wc[w].sum():.0f}")
        # correlation: warp time vs max-lane niter and coupled
        import numpy as np
        print("    corr(kcycles, max niter) %.2f  corr(kcycles, coupled) %.2f" % (np.corrcoef(wk.numpy(), wn.max(dim=1).values.numpy())[0,1], np.corrcoef(wk.numpy(), wc.sum(dim=1).numpy())[0,1]))
