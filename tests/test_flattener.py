"""CPU: the MJCF flattener -- sizes and topology of the three scenes, and, where the reference
snapshot is mounted (this container, not the GPU box), that the committed flat models and scene headers are exactly
what the flattener derives from the reference XML today."""
import json
import os

import numpy as np
import pytest

from gym_kmanip_b200 import mjcf, scenegen

REF_ASSETS = "/root/reference/gym_kmanip/assets"


@pytest.mark.parametrize("scene,nbody,njnt,nq,nv,nu", [("solo_arm", 17, 11, 17, 16, 10), ("dual_arm", 30, 21, 27, 26, 20),
                                                        ("torso", 29, 21, 27, 26, 20)])
def test_scene_sizes(scene, nbody, njnt, nq, nv, nu):
    f = mjcf.load_flat(scene)
    assert (f["nbody"], f["njnt"], f["nq"], f["nv"], f["nu"]) == (nbody, njnt, nq, nv, nu)
    assert f["opt"]["timestep"] == 0.002            # MuJoCo default: no <option> in the reference assets
    assert f["jnt_type"][-1] == mjcf.JNT_FREE        # the cube owns the last 7 qpos / 6 dofs
    par = f["body_parent"]
    assert all(par[b] < b for b in range(1, nbody))  # parents before children
    depth = [0] * nbody
    for b in range(1, nbody):
        depth[b] = depth[par[b]] + 1
    assert max(depth) == 10
    M0 = np.array(f["dof_invweight0"])
    assert (M0 > 0).all() and f["meaninertia"] > 0
    # position servos of arm_r.xml:46-55 / torso.xml:113-134
    kp = sorted(set(f["act_kp"]))
    assert kp == ([0.0, 200.0, 1000.0] if scene != "torso" else [100.0])
    assert f["npair"] >= 3 and f["geom_name"][-2:] == ["table", "cube"]


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference snapshot not mounted (GPU box)")
@pytest.mark.parametrize("scene", ["solo_arm", "dual_arm", "torso"])
def test_committed_flat_models_are_current(scene):
    flat = mjcf.flatten(os.path.join(REF_ASSETS, mjcf.SCENE_FILES[scene]))
    committed = mjcf.load_flat(scene)
    a, b = json.loads(json.dumps(flat)), json.loads(json.dumps(committed))
    assert a.keys() == b.keys()
    for key in a:
        if isinstance(a[key], (list, float)) and key not in ("body_name", "jnt_name", "site_name", "geom_name", "cam_name"):
            assert np.allclose(np.array(a[key], dtype=float), np.array(b[key], dtype=float), rtol=1e-13, atol=1e-15), key
        else:
            assert a[key] == b[key], key
    hdr = os.path.join(os.path.dirname(mjcf.__file__), "csrc", "scenes", f"scene_{scene}.h")
    assert scenegen.render_scene_header(scene, flat) == open(hdr).read()
