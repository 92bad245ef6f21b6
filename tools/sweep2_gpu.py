"""Mapping sweep: env-steps/s of the warp-per-env (lanes 32) and thread-per-env (lanes 2) kernels per scene / batch size."""
import sys, os, json, subprocess
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = [("KManipSoloArmQPos", "float32", [8192, 12288, 16384, 32768, 65536]), ("KManipSoloArm", "float32", [8192, 16384]),
         ("KManipDualArm", "float32", [4096, 8192, 32768]), ("KManipTorso", "float64", [4096, 8192, 16384]), ("KManipTorso", "float32", [8192, 32768])]
for env, dt, ns in cases:
    for n in ns:
        row = []
        for lanes in (32, 2):
            out = subprocess.run([sys.executable, os.path.join(R, "bench.py"), "--env", env, "--dtype", dt, "--envs", str(n), "--lanes", str(lanes),
                                  "--steps", "8", "--warmup", "24", "--no-cpu", "--no-extra"], capture_output=True, text=True, timeout=120).stdout
            try:
                d = json.loads(out.strip().splitlines()[-1]); row.append(f"lanes {lanes}: {d['value'] / 1e6:.2f}e6 ({d['ms_per_step']:.2f} ms)")
            except Exception as e:
                row.append(f"lanes {lanes}: failed")
        print(env, dt, n, " | ".join(row), flush=True)
