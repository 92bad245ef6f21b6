"""Quick check + timing of launch configurations (lanes per env G = 32 / 16 / 1) against each other."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from gym_kmanip_b200.batch_sim import BatchSim

def run(env, n, G, epb, steps=20, dtype="float32"):
    sim = BatchSim(env, n, dtype=dtype, seed=1)
    cfg = sim.configure(G, epb)
    sim.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.rand(8, n, sim.act_dim, device="cuda", generator=gen) * 2 - 1
    for i in range(24): sim.step(acts[i % 8], contacts=False)     # into the contact-rich part of the episode
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): sim.step(acts[i % 8], contacts=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = sim.get_state()[0].double().cpu()
    sim.close()
    return ms, cfg, st

if __name__ == "__main__":
    envs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["KManipSoloArmQPos"]
    for env in envs:
        for n in (4096, 16384, 65536):
            ref = None
            for G, epb in ((32, 0), (1, 32), (1, 64)):
                try:
                    ms, cfg, st = run(env, n, G, epb)
                except Exception as ex:
                    print(env, n, G, epb, "FAILED", ex); continue
                d = "" if ref is None else " max|dstate| vs G=32: %.2e" % float((st - ref).abs().max())
                if ref is None: ref = st
                print(f"{env} n={n} G={G} epb={cfg['envs_per_block']} grid={cfg['grid']} ctas/sm={cfg['ctas_per_sm']}: {ms:.3f} ms/step {n/ms*1e3:.3e} env-steps/s{d}", flush=True)
