"""GPU: the reference-side binding printed in INTEGRATION.md section 2 is executed as written (only the library path is
made absolute and `gym_kmanip` is stood in for by this package's constants) and must behave like the shipped backend:
same observations / rewards as gym_kmanip_b200.env_sim on the same seed and actions, and a camera image from k_render."""
import os
import re
import sys
import types

import numpy as np
import pytest

import gym_kmanip_b200 as k
from gym_kmanip_b200 import _lib, constants as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _documented_binding():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# gym_kmanip/env_b200\.py.*?)```", text, re.S).group(1)
    return block.replace('"libkmanip_b200.so"', repr(_lib.LIB_PATH))


def test_documented_binding_is_present_and_parses():
    src = _documented_binding()
    compile(src, "INTEGRATION.md:env_b200", "exec")
    for name in ("km_create", "km_reset_host", "km_step_host", "km_render_host", "km_destroy"):
        assert name in src and name in _lib.EXPORTS


@pytest.mark.gpu
def test_documented_binding_runs_like_the_shipped_backend():
    stub = types.ModuleType("gym_kmanip")
    stub.CONTROL_TIMESTEP = K.CONTROL_TIMESTEP
    sys.modules["gym_kmanip"] = stub
    try:
        ns = {}
        exec(compile(_documented_binding(), "INTEGRATION.md:env_b200", "exec"), ns)
        kw = K.ENV_REGISTRY["KManipSoloArm"]
        gym_env = types.SimpleNamespace(seed=3, q_len=len(kw["q_pos_home"]), **{f: kw.get(f) for f in (
            "mjcf_filename", "q_pos_home", "q_id_r_mask", "q_id_l_mask", "ctrl_id_r_grip", "ctrl_id_l_grip", "obs_list", "act_list")})
        b = ns["new"](gym_env)
        ref = k.make("KManipSoloArm", seed=3, ik_mode=0).unwrapped       # shipped backend, fp64, same IK mode as km_create's task default
        term, rew, disc, obs, t = b.k_reset()
        robs, _ = ref.reset()
        assert term is False and rew is None and t == 0.0 and list(obs) == kw["obs_list"]
        for key in obs:
            assert np.allclose(obs[key], robs[key], atol=1e-12), key
        rng = np.random.default_rng(0)
        for _ in range(3):
            act = {"eer_pos": rng.uniform(-1, 1, 3).astype(np.float32), "eer_orn": rng.uniform(-1, 1, 3).astype(np.float32),
                   "grip_r": rng.uniform(-1, 1, 1).astype(np.float32)}
            term, rew, disc, obs, t = b.k_step(act)
            robs, rrew, *_ = ref.step(act)
            assert abs(rew - rrew) < 1e-12 and all(np.allclose(obs[key], robs[key], atol=1e-12) for key in obs)
        img = b.k_render(K.CAMERAS["top"])
        assert img.shape == (480, 640, 3) and img.dtype == np.uint8 and np.array_equal(img, ref.render())
        b.k_close()
        ref.close()
    finally:
        del sys.modules["gym_kmanip"]
