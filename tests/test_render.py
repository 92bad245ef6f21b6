"""Camera observations of the Vision ids: reference env_sim.py:140-145 / :187-188.

CPU: the numpy restatement (oracle/render_oracle.py) against its committed golden images, camera conventions, and the
host-side camera / light structs.  GPU: km_render (csrc/km_render.cuh) against the oracle -- the render records of the
setup kernel against the oracle's (from the C++ oracle's body frames), the pixel kernel against the numpy ray caster
on the SAME records at the reference's full camera sizes (tile culling included), end to end against the golden images,
and the Vision ids through the reference-facing classes.  Images are uint8; a pixel may differ by one level where a
float32 product rounds differently, and silhouette pixels may flip -- the bounds are written at each assertion."""
import os

import numpy as np
import pytest

import gym_kmanip_b200 as k
from gym_kmanip_b200 import constants as K, mjcf, render as R
from oracle import render_oracle as ro

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RENDER_ENVS = {"KManipSoloArm": ["head", "grip_r"], "KManipDualArm": ["head", "grip_l", "grip_r"], "KManipTorso": ["head", "grip_l", "grip_r"]}
RENDER_SIZE = {"head": (160, 120), "grip_r": (60, 40), "grip_l": (60, 40)}


def _flat(env_id):
    return mjcf.load_flat(mjcf.scene_of_mjcf(K.ENV_REGISTRY[env_id]["mjcf_filename"]))


def image_diff(a, b):
    """(fraction of pixels off by more than one level in some channel, largest channel difference)"""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32)).max(axis=-1)
    return float((d > 1).mean()), int(d.max())


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("env_id", list(RENDER_ENVS))
def test_oracle_reproduces_golden_images(env_id):
    from oracle.oracle import Oracle
    g = np.load(os.path.join(GOLD, f"render_{env_id}.npz"))
    o = Oracle(env_id)
    st = g["state"][1]
    o.set_state(st[:o.nq], st[o.nq:o.nq + o.nv], st[o.nq + o.nv:o.nq + o.nv + o.nu])
    o.mj_forward(disable_actuation=True)
    xp, xq = o.field("xpos"), o.field("xquat")
    for cam in RENDER_ENVS[env_id]:
        w, h = RENDER_SIZE[cam]
        rec = ro.scene_record(o.flat, xp, xq, cam)
        assert np.array_equal(rec.view(np.int32), g[f"rec_{cam}"][1].view(np.int32))
        assert np.array_equal(ro.render_record(rec, ro.params(o.flat, cam, w, h)), g[f"img_{cam}"][1])


def test_camera_conventions():
    """mj_camlight targetbody frame, pinhole projection and what the head camera must see at the home pose."""
    from oracle.oracle import Oracle
    o = Oracle("KManipSoloArm")
    o.reset(np.array([0.2, 0.6, 0.65]))
    o.mj_forward(disable_actuation=True)
    xp, xq = o.field("xpos").reshape(-1, 3), o.field("xquat").reshape(-1, 4)
    rec = ro.scene_record(o.flat, xp, xq, "head")
    org, X, Y, Z = rec[0:3], rec[3:6], rec[6:9], rec[9:12]
    assert np.allclose(org, [0, 0, 1.0]) and np.allclose(Z, ro._unit(np.array([0, 0, 1.0]) - np.array([0, 0.6, 0.5])), atol=1e-6)
    assert abs(X[2]) < 1e-7 and np.allclose(np.cross(Z, X), Y, atol=1e-6) and Y[2] > 0       # x horizontal, y up
    assert abs(np.dot(X, Z)) < 1e-6 and np.isclose(np.linalg.norm(X), 1, atol=1e-6)
    cam = K.CAMERAS["head"]
    P = ro.params(o.flat, "head", cam.w, cam.h)
    assert np.isclose(P["focal"], 0.5 * 480 / np.tan(np.deg2rad(39.0)), rtol=1e-6) and np.isclose(P["tab_z"], 0.5)
    img = ro.render_record(rec, P)
    assert img.shape == (480, 640, 3) and img.dtype == np.uint8
    # the cube (red) projects where the pinhole model says: u = W/2 + f x/(-z), v = H/2 - f y/(-z) in the camera frame
    pc = np.array([0.2, 0.6, 0.65]) - org
    u, v = 320 + P["focal"] * np.dot(pc, X) / -np.dot(pc, Z), 240 - P["focal"] * np.dot(pc, Y) / -np.dot(pc, Z)
    red = (img[..., 0] > 150) & (img[..., 1] < 60) & (img[..., 2] < 60)
    ys, xs = np.nonzero(red)
    assert red.sum() > 100 and abs(xs.mean() - u) < 4 and abs(ys.mean() - v) < 4
    # the table (rgba 0.2, lit) fills the frame away from the robot; no pixel is background
    assert (img.sum(-1) > 0).all() and 40 < img[470, 10, 0] < 120 and img[470, 10, 0] == img[470, 10, 1] == img[470, 10, 2]


def test_host_camera_structs():
    for env_id, body in (("KManipSoloArm", "hand"), ("KManipTorso", "hand")):
        flat = _flat(env_id)
        c = R.camera_struct(flat, "head", 640, 480)
        assert c.link == -1 and list(c.pos) == [0, 0, 1.0] and c.target_link == -1 and np.allclose(list(c.target_pos), [0, 0.6, 0.5])
        assert c.fovy == 78.0 and (c.width, c.height) == (640, 480)
        g = R.camera_struct(flat, "grip_r", 60, 40)
        # the gripper camera and the body it tracks (eer_site) ride on the same hand link; static bodies are folded
        assert g.link >= 0 and g.link == g.target_link and g.fovy == 20.0
        ci = flat["cam_name"].index("grip_r")
        assert flat["jnt_bodyid"][g.link] == flat["cam_bodyid"][ci] and np.allclose(list(g.pos), flat["cam_pos"][ci])
        tb = flat["cam_targetbodyid"][ci]
        assert flat["body_parent"][tb] == flat["cam_bodyid"][ci] and np.allclose(list(g.target_pos), flat["body_pos"][tb])
        v = R.visual_struct(flat)
        assert v.nlight == 3 and np.allclose(list(v.head_ambient), 0.4) and np.allclose(list(v.rgb_cube), [1, 0, 0])
        assert np.allclose(list(v.rgb_table), 0.2) and np.isclose(np.linalg.norm(list(v.light_dir[0])), 1.0)
    with pytest.raises(KeyError):
        R.camera_struct(_flat("KManipSoloArm"), "grip_l", 60, 40)


# ------------------------------------------------------------------------------------------------ GPU
def _sim_with_states(env_id, states, dtype):
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    s = BatchSim(env_id, len(states), dtype=dtype, seed=0)
    s.reset()
    s.set_state(torch.as_tensor(states))
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", list(RENDER_ENVS))
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_render_matches_oracle(env_id, dtype):
    g = np.load(os.path.join(GOLD, f"render_{env_id}.npz"))
    s = _sim_with_states(env_id, g["state"], dtype)
    flat = s.flat
    for cam in RENDER_ENVS[env_id]:
        w, h = RENDER_SIZE[cam]
        c = K.Cam(w, h, 3, 0, (w // 2, h // 2), cam, f"camera/{cam}")
        img = s.render(c).cpu().numpy()
        rec = s.render_records().cpu().numpy()
        ref = g[f"rec_{cam}"]
        # setup kernel: same primitive kinds / materials bit for bit, frames to the precision of the build
        assert np.array_equal(rec[:, 12].view(np.int32), ref[:, 12].view(np.int32))
        assert np.array_equal(rec[:, 16::16].view(np.int32), ref[:, 16::16].view(np.int32))
        body = np.delete(rec, np.r_[12, 16:rec.shape[1]:16], axis=1), np.delete(ref, np.r_[12, 16:rec.shape[1]:16], axis=1)
        assert np.abs(body[0] - body[1]).max() < (2e-7 if dtype == "float64" else 5e-6)
        # pixel kernel on its own records against the numpy ray caster: at most 1 level, a few silhouette pixels aside
        P = ro.params(flat, cam, w, h)
        for i in range(len(rec)):
            bad, _ = image_diff(img[i], ro.render_record(rec[i], P))
            assert bad <= 2e-3, (cam, i, bad)
        # end to end against the committed golden images (oracle body frames in fp64)
        bad, _ = image_diff(img, g[f"img_{cam}"])
        assert bad <= (2e-3 if dtype == "float64" else 5e-3), (cam, bad)
    s.close()


@pytest.mark.gpu
def test_render_full_size_cameras_and_culling():
    """The reference's camera sizes (640 x 480 head / top, 60 x 40 gripper): every tile of the pixel kernel, culled
    primitive lists included, against the un-culled numpy ray caster on the same records."""
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    s = BatchSim("KManipDualArm", 3, dtype="float32", seed=5)
    s.reset()
    gen = torch.Generator(device="cuda").manual_seed(2)
    for _ in range(12):
        s.step(torch.rand(3, s.act_dim, device="cuda", generator=gen) * 2 - 1)
    for name in ("head", "top", "grip_l"):
        cam = K.CAMERAS[name]
        img = s.render(cam).cpu().numpy()
        rec = s.render_records().cpu().numpy()
        assert img.shape == (3, cam.h, cam.w, 3)
        P = ro.params(s.flat, name, cam.w, cam.h)
        for i in range(3):
            bad, worst = image_diff(img[i], ro.render_record(rec[i], P))
            assert bad <= 1e-3, (name, i, bad, worst)
    host = np.zeros((3, 40, 60, 3), dtype=np.uint8)
    s.render_host("grip_r", host)
    assert np.array_equal(host, s.render("grip_r").cpu().numpy())
    s.close()


@pytest.mark.gpu
def test_vision_ids_through_the_reference_api():
    import torch
    env = k.make("KManipSoloArmVision", ik_mode=0)
    u = env.unwrapped
    obs, info = env.reset(seed=0)
    assert list(obs) == ["q_pos", "q_vel", "camera/head", "camera/grip_r"] and [c.name for c in info["cameras"]] == ["head", "grip_r"]
    assert obs["camera/head"].shape == (480, 640, 3) and obs["camera/grip_r"].shape == (40, 60, 3) and obs["camera/head"].dtype == np.uint8
    assert u.observation_space.contains(obs)
    u.action_space.seed(0)
    obs2, *_ = env.step(u.action_space.sample())
    assert obs2["camera/head"].any() and u.render().shape == (480, 640, 3)          # render(): the "top" camera (env_base.py:216-217)
    env.close()
    from gym_kmanip_b200.vector_env import KManipVectorEnv
    venv = KManipVectorEnv("KManipDualArmVision", 5, seed=1)
    vobs, _ = venv.reset()
    assert vobs["camera/head"].shape == (5, 480, 640, 3) and vobs["camera/grip_l"].shape == (5, 40, 60, 3) and vobs["camera/head"].is_cuda
    before = vobs["camera/grip_r"].clone()
    for _ in range(3):
        vobs, rew, term, trunc, vinfo = venv.step(venv.sample_actions())
    assert not torch.equal(before, vobs["camera/grip_r"]) and "camera/head" not in vinfo["final_obs"]
    # envs render independently: env 2 of the batch equals the same state rendered alone
    st, _, _ = venv.sim.get_state()
    from gym_kmanip_b200.batch_sim import BatchSim
    one = BatchSim("KManipDualArmVision", 1, seed=1)
    one.reset()
    one.set_state(st[2:3])
    assert torch.equal(one.render("head")[0], vobs["camera/head"][2])
    one.close()
    venv.close()
