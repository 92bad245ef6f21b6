"""Per-phase cycle profile of the env-step kernel (needs a library built with -DKM_PHASE_CLOCKS, see csrc/Makefile `dbg`).
usage: KMANIP_B200_LIB=gym_kmanip_b200/lib_dbg/libkmanip_b200.so python tools/phase_clocks_gpu.py [env] [n] [lanes] [epb]"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from gym_kmanip_b200 import _lib
from gym_kmanip_b200.batch_sim import BatchSim
env = sys.argv[1] if len(sys.argv) > 1 else "KManipSoloArmQPos"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
epb = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ik_mode = int(os.environ.get("IK_MODE", "0"))
sim = BatchSim(env, n, dtype="float32", seed=0, ik_mode=ik_mode)
if lanes or epb:
    sim.configure(lanes, epb)
print("launch", sim.launch_config())
clk = torch.zeros(n, 16, dtype=torch.int32, device="cuda")
_lib.check(sim.L.km_debug_phase_clocks(sim.h, clk.data_ptr()))
sim.reset()
gen = torch.Generator(device="cuda").manual_seed(1234)
names = ["kin", "crb", "coll+mkc", "vel", "acc", "sol_setup", "sol_dir", "sol_ls", "sol_upd", "sol_vote", "euler", "barrier", "before", "epilogue"]
for t in range(40):
    act = torch.rand(n, sim.act_dim, device="cuda", generator=gen) * 2 - 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.step(act, contacts=True); e1.record(); torch.cuda.synchronize()
    if t in (3, 8, 15, 30, 39):
        c = clk.cpu().double()
        tot = c.sum(1)
        print(f"step {t}: {e0.elapsed_time(e1):.3f} ms; cycles per env step (lane 0 of each env): mean total {tot.mean():.0f} max {tot.max():.0f}")
        print("   " + "  ".join(f"{nm} {c[:, i].mean() / tot.mean() * 100:.1f}%" for i, nm in enumerate(names)))
        # per-CTA view (envs of a CTA march in phase, so their totals agree): which CTAs set the kernel time, and why
        epb_ = sim.launch_config()["envs_per_block"]
        ncta = (n + epb_ - 1) // epb_
        pad = ncta * epb_ - n
        tp = torch.cat([tot, tot.new_zeros(pad)]).view(ncta, epb_).max(1).values
        q = torch.quantile(tp, torch.tensor([0.1, 0.5, 0.9, 0.99, 1.0], dtype=tp.dtype))
        print("   per-CTA total cycles p10/p50/p90/p99/max: " + " ".join(f"{x:.0f}" for x in q.tolist()))
        busy = c[:, :9].sum(1) + c[:, 10] + c[:, 12] + c[:, 13]          # everything but the barrier waits
        bq = torch.quantile(busy, torch.tensor([0.1, 0.5, 0.9, 0.99, 1.0], dtype=tp.dtype))
        print("   per-env busy cycles (no barrier waits) p10/p50/p90/p99/max: " + " ".join(f"{x:.0f}" for x in bq.tolist()))
        bf = torch.quantile(c[:, 12], torch.tensor([0.1, 0.5, 0.9, 0.99, 1.0], dtype=tp.dtype))
        print("   per-env before_step (action decode + IK) cycles p10/p50/p90/p99/max: " + " ".join(f"{x:.0f}" for x in bf.tolist()))
        top = torch.argsort(busy, descending=True)[:6]
        fl = sim.con_flags.cpu(); nc = sim.ncon.cpu()
        for i in top.tolist():
            print(f"     env {i} (cta {i // epb_}) busy {busy[i]:.0f} ncon {int(nc[i])} flags {int(fl[i])} | " + " ".join(f"{nm} {c[i, k]:.0f}" for k, nm in enumerate(names)) + f" | newton its {c[i, 9]:.0f} cube rebuilds {c[i, 14]:.0f} chain refactors {c[i, 15]:.0f}")
        print(f"   batch means: newton iterations {c[:, 9].mean():.1f}, cube-block rebuilds {c[:, 14].mean():.1f}, chain-block refactors {c[:, 15].mean():.1f} per env step")
