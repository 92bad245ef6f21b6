"""Launch-configuration sweep (lanes per env, envs per CTA) at fixed batch sizes."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
from tpe_gpu import run
env = sys.argv[1] if len(sys.argv) > 1 else "KManipSoloArmQPos"
ns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4096, 16384]
cfgs = [tuple(int(y) for y in x.split(":")) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [(32, 0), (32, 14), (32, 9), (32, 7), (32, 4), (16, 0), (16, 28), (16, 14), (16, 8)]
for n in ns:
    for G, epb in cfgs:
        try:
            ms, cfg, st = run(env, n, G, epb)
            print(f"{env} n={n} G={G} epb={cfg['envs_per_block']} grid={cfg['grid']} ctas/sm={cfg['ctas_per_sm']} smem={cfg['smem_bytes']}: {ms:.3f} ms/step {n/ms*1e3:.3e} env-steps/s", flush=True)
        except Exception as ex:
            print(env, n, G, epb, "FAILED", str(ex)[:100], flush=True)
