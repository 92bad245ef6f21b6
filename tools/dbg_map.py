import sys, os
sys.path.insert(0, os.getcwd())
import torch
from gym_kmanip_b200.batch_sim import BatchSim
for n in (512, 4096):
    outs = {}
    for lanes in (32, 2, 1):
        s = BatchSim("KManipSoloArm", n, dtype="float32", seed=4)
        cfg = s.configure(lanes, 0)
        s.reset()
        gen = torch.Generator(device="cuda").manual_seed(1)
        act = torch.rand(n, s.act_dim, device="cuda", generator=gen) * 2 - 1
        o = s.step(act, autoreset=True)[0].clone()
        torch.cuda.synchronize()
        outs[lanes] = o
        print(n, lanes, cfg)
        s.close()
    for lanes in (2, 1):
        d = (outs[lanes] - outs[32]).abs().max(dim=1).values
        bad = (d > 1e-3).nonzero().flatten()
        print(" n", n, "lanes", lanes, "max diff %.3e" % float(d.max()), "bad rows", bad.numel(), bad[:20].tolist())
