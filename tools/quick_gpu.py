import sys, time
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from gym_kmanip_b200.batch_sim import BatchSim
for env, n in (("KManipSoloArmQPos", 4096), ("KManipSoloArm", 8192), ("KManipDualArm", 8192)):
    for G in (32,):
        sim = BatchSim(env, n, dtype="float32")
        try:
            cfg = sim.configure(G, 0)
        except Exception as e:
            print(env, G, "cfg fail", e); continue
        sim.reset()
        act = torch.rand(n, sim.act_dim, device="cuda") * 2 - 1
        for _ in range(3): sim.step(act)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 20
        for _ in range(K): sim.step(act)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(env, n, cfg, "ms/step %.3f" % ms, "env-steps/s %.3e" % (n / ms * 1e3), flush=True)
        sim.close()
