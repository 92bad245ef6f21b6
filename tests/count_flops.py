#!/usr/bin/env python
"""Algorithmic FLOPs per env step, counted (not estimated) by the oracle's operation-counting build
(oracle/kmanip_oracle.cpp compiled with -DKO_COUNT_FLOPS: every floating-point +,-,*,/ and sqrt of the scalar type
increments a counter; transcendental calls count as one).  Writes profiles/flops_per_env_step.json, which bench.py
uses for the FP-pipe roofline.  The workload is the bench workload: random actions, autoreset every 64 steps,
averaged over one full episode of 64 envs."""
import json, os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)   # tests/ -> repo root
from oracle import oracle as om

out = {}
for env_id in ["KManipSoloArmQPos", "KManipSoloArm", "KManipDualArm", "KManipDualArmQPos", "KManipTorso"]:
    o = om.Oracle(env_id, flops=True)
    n, steps = 64, 64
    st = om.batch_reset_state(o, n, seed=0)
    if o.nmocap == 0:
        st["mocap"] = np.zeros((n, 0))
    rng = np.random.default_rng(0)
    o.flops_reset()
    per_step = []
    for t in range(steps):
        a = rng.uniform(-1, 1, (n, o.task.act_dim)).astype(np.float32)
        f0 = o.flops()
        om.batch_step(o, st, a, autoreset=True, seed=0, nthreads=1)
        per_step.append((o.flops() - f0) / n)
    out[env_id] = float(np.mean(per_step))
    print(env_id, "FLOP/env-step: mean %.3e  first-steps (free flight) %.3e  late (cube resting) %.3e" % (np.mean(per_step), np.mean(per_step[:5]), np.mean(per_step[-20:])))
out["note"] = "counted by oracle/kmanip_oracle.cpp -DKO_COUNT_FLOPS over one 64-step episode of 64 envs (tests/count_flops.py); an fma counts as 2"
path = os.path.join(R, "profiles", "flops_per_env_step.json")
old = {}
if os.path.exists(path):
    old = json.load(open(path))
old.update(out)
json.dump(old, open(path, "w"), indent=1)
print("wrote", path)
