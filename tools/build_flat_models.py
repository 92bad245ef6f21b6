#!/usr/bin/env python
"""Generate the committed flat models and the CUDA scene headers from the reference MJCF.

usage: python tools/build_flat_models.py [/root/reference/gym_kmanip/assets]

Reads the reference XML where it lies (read-only), applies gym_kmanip_b200/assets/completion_spec.json
and writes
  gym_kmanip_b200/assets/flat/{solo_arm,dual_arm,torso}.json   (derived numbers, not reference sources)
  gym_kmanip_b200/csrc/scenes/scene_{...}.h                     (compile-time topology for the kernels)
Nothing at run time reads /root/reference.
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gym_kmanip_b200 import mjcf  # noqa: E402
from gym_kmanip_b200 import scenegen  # noqa: E402


def main():
    assets = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/gym_kmanip/assets"
    for scene, fname in mjcf.SCENE_FILES.items():
        flat = mjcf.flatten(os.path.join(assets, fname))
        out = os.path.join(mjcf.FLAT_DIR, f"{scene}.json")
        with open(out, "w") as f:
            json.dump(flat, f, indent=1)
        hdr = scenegen.write_scene_header(scene, flat)
        print(f"{scene}: nbody={flat['nbody']} njnt={flat['njnt']} nq={flat['nq']} nv={flat['nv']} "
              f"nu={flat['nu']} ngeom={flat['ngeom']} npair={flat['npair']} -> {out}, {hdr}")


if __name__ == "__main__":
    main()
