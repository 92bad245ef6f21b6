"""Times km_render (camera observations) on the GPU: images/s and achieved HBM write bandwidth of the pixel kernel.

usage: python tools/render_bench_gpu.py [--env KManipSoloArmVision] [--envs 4096] [--cam head] [--iters 20]
Algorithmic bytes per launch = n * h * w * 3 (the image batch) + n * record bytes; CUDA events on the launching stream.
"""
import argparse
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch  # noqa: E402

from gym_kmanip_b200 import constants as K  # noqa: E402
from gym_kmanip_b200.batch_sim import BatchSim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="KManipSoloArmVision")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--cam", default="head")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    s = BatchSim(a.env, a.envs, dtype="float32", seed=0)
    s.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(8):
        s.step(torch.rand(a.envs, s.act_dim, device="cuda", generator=g) * 2 - 1)
    cam = K.CAMERAS[a.cam]
    out = torch.empty(a.envs, cam.h, cam.w, 3, dtype=torch.uint8, device="cuda")
    for _ in range(a.warmup):
        s.render(cam, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
    ev[0].record()
    for i in range(a.iters):
        s.render(cam, out=out)          # the image batch (3.8 GB at 4096 x 640 x 480) is far larger than L2
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters))
    med = ms[len(ms) // 2]
    nbytes = out.numel() + a.envs * int(s.L.km_render_record_floats(s.h)) * 4
    peak = None
    try:
        peak = json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    line = dict(kernel="km_render (k_render_setup + k_render_pixels)", env=a.env, envs=a.envs, camera=a.cam, width=cam.w, height=cam.h,
                ms_per_launch=med, ms_min=ms[0], images_per_s=a.envs / (med * 1e-3), algorithmic_bytes=nbytes,
                achieved_GBps=nbytes / (med * 1e-3) / 1e9, mean_pixel=float(out[:8].float().mean()))
    if peak:
        for kx in ("hbm_gbs",):
            if kx in peak:
                line["peak_GBps"] = peak[kx]
                line["frac"] = line["achieved_GBps"] / peak[kx]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
