"""CPU: host-side logic of the drop-in boundary -- registry, spaces, flat action / observation layouts, task structs,
env sharding -- mirroring what gymnasium's check_env pins for the reference (tests/test_env.py:8-24)."""
import json
import os

import numpy as np
import pytest

import gym_kmanip_b200 as k
from gym_kmanip_b200 import constants as K, flatmodel, mjcf, sharding
from gym_kmanip_b200.spaces import Box, Dict

STATE_IDS = ["KManipSoloArm", "KManipSoloArmQPos", "KManipDualArm", "KManipDualArmQPos", "KManipTorso"]


def test_registry_has_the_eight_reference_ids():
    assert sorted(K.ENV_REGISTRY) == sorted(["KManipSoloArm", "KManipSoloArmQPos", "KManipSoloArmVision", "KManipDualArm",
                                             "KManipDualArmQPos", "KManipDualArmVision", "KManipTorso", "KManipTorsoVision"])
    assert k.MAX_EPISODE_STEPS == 64 and k.CONTROL_TIMESTEP == 0.02 and k.OBS_DTYPE is np.float64 and k.ACT_DTYPE is np.float32
    for kw in K.ENV_REGISTRY.values():
        assert len(kw["q_keys"]) == len(kw["q_pos_home"]) == len(kw["q_dict"])
        assert kw["q_pos_home"].dtype == np.float32


@pytest.mark.parametrize("env_id,act_dim,obs_dim", [("KManipSoloArm", 7, 27), ("KManipSoloArmQPos", 8, 27),
                                                     ("KManipDualArm", 14, 47), ("KManipDualArmQPos", 16, 47),
                                                     ("KManipTorso", 14, 47)])
def test_flat_layouts(env_id, act_dim, obs_dim):
    kw = K.ENV_REGISTRY[env_id]
    flat = mjcf.load_flat(mjcf.scene_of_mjcf(kw["mjcf_filename"]))
    t = flatmodel.make_task(flat, kw)
    assert t.act_dim == act_dim and 2 * t.q_len + 7 == obs_dim
    n_r = len(kw["q_id_r_mask"])
    n_l = len(kw["q_id_l_mask"]) if kw.get("q_id_l_mask") is not None else 0
    lay = flatmodel.action_layout(kw["act_list"], n_r, n_l)
    # keys in the order env_base.py:149-190 inserts them; contiguous, non-overlapping
    assert list(lay) == [x for x in K.ACTION_KEY_ORDER if x in kw["act_list"]]
    o = 0
    for sl in lay.values():
        assert sl.start == o
        o = sl.stop
    assert o == act_dim
    ol = flatmodel.obs_layout(t.q_len)
    assert list(ol) == ["q_pos", "q_vel", "cube_pos", "cube_orn"] and ol["cube_orn"].stop == obs_dim
    # the right arm is processed first (env_sim.py:60-99)
    assert t.n_arm == (1 if "Solo" in env_id else 2)
    assert [t.arm_mask[0][i] for i in range(t.arm_nmask[0])] == list(kw["q_id_r_mask"])
    assert t.act_mode == (1 if "QPos" in env_id else 0)
    assert t.max_episode_steps == 64 and t.ik_teleport == 1


def test_spaces_shim_contract():
    b = Box(-1, 1, (3,), np.float32, seed=0) if Box.__module__.endswith("spaces") else Box(-1, 1, (3,), np.float32)
    x = b.sample()
    assert x.shape == (3,) and x.dtype == np.float32 and b.contains(x)
    assert not b.contains(np.array([2, 0, 0], dtype=np.float32))
    d = Dict({"a": Box(-1, 1, (2,), np.float64), "b": Box(-1, 1, (1,), np.float32)})
    s = d.sample()
    assert list(s) == ["a", "b"] and d.contains(s)


def test_shard_ranges_partition_the_job():
    for total in (0, 1, 7, 4096, 65536, 100003):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                e0, n = sharding.shard_range(total, r, world)
                cover += list(range(e0, e0 + n)) if total < 5000 else []
                assert n in (total // world, total // world + 1)
            if total < 5000:
                assert cover == list(range(total))
                for g in range(0, total, max(1, total // 50)):
                    r, i = sharding.owner_of(g, total, world)
                    e0, n = sharding.shard_range(total, r, world)
                    assert e0 + i == g and 0 <= i < n
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_rerun_logger_and_real_robot_are_declared_out_of_scope():
    from gym_kmanip_b200.env_base import KManipEnv
    with pytest.raises(NotImplementedError):
        KManipEnv(**dict(K.ENV_REGISTRY["KManipSoloArm"], sim=False))
    with pytest.raises(NotImplementedError):
        KManipEnv(**dict(K.ENV_REGISTRY["KManipSoloArm"], log_rerun=True))
    with pytest.raises(KeyError):
        k.make("KManipNope")


def test_episode_log_layout_follows_the_reference(tmp_path):
    """log_episode writes the reference's ACT layout (log_h5py.py:13-61), quirks included: the logged qpos / qvel are
    the normalised observations and every action row is grip_r broadcast over a_len columns."""
    from gym_kmanip_b200 import log_episode
    info = {"step": 0, "episode": 3, "is_success": False, "q_keys": ["a", "b"], "q_len": 10, "a_len": 3, "obs_list": ["q_pos"],
            "act_list": ["eer_pos", "eer_orn", "grip_r"], "cameras": [], "sim": True, "cpu_time": 1.0, "reward": None}
    f = log_episode.new(str(tmp_path), info)
    rng = np.random.default_rng(0)
    rows = []
    for t in range(1, 6):
        info["step"] = t
        act = {"eer_pos": rng.uniform(-1, 1, 3).astype(np.float32), "eer_orn": rng.uniform(-1, 1, 3).astype(np.float32),
               "grip_r": rng.uniform(-1, 1, 1).astype(np.float32)}
        obs = {"q_pos": rng.uniform(0, 1, 10), "q_vel": rng.uniform(-1, 1, 10)}
        log_episode.step(f, act, obs, info)
        rows.append((act, obs))
    path = log_episode.end(f)
    assert os.path.basename(path).startswith("episode_3.")
    assert path.endswith(".hdf5")
    qpos, qvel, action, attrs, meta = log_episode.read_episode(path)
    assert attrs["sim"] and meta["q_len"] == 10 and meta["episode"] == 3
    assert qpos.shape == (64, 10) and qvel.shape == (64, 10) and action.shape == (64, 3)
    for t, (act, obs) in enumerate(rows):
        assert np.allclose(qpos[t], obs["q_pos"], atol=1e-7) and np.allclose(qvel[t], obs["q_vel"], atol=1e-7)
        assert np.all(action[t] == act["grip_r"][0])
    assert not qpos[5:].any()


def test_batch_episode_log_ring_buffers(tmp_path):
    """BatchEpisodeLog: ring buffers (here CPU tensors; CUDA tensors in the batched env) -> one file per finished
    episode of each logged env, same layout and quirks as the single-env logger; the finished episode's last row is the
    final observation, not the first observation of the next episode."""
    import torch
    from gym_kmanip_b200.log_episode import BatchEpisodeLog
    n, q_len, act_dim, T = 6, 10, 7, 4
    log = BatchEpisodeLog(str(tmp_path), [1, 4], q_len, 3, 6, torch, "cpu", env0=100, max_steps=64, info={"env_id": "KManipSoloArm"})
    g = torch.Generator().manual_seed(0)
    hist = []
    for t in range(2 * T + 1):
        act = torch.rand(n, act_dim, generator=g) * 2 - 1
        obs, fin = torch.rand(n, 27, generator=g, dtype=torch.float64), torch.rand(n, 27, generator=g, dtype=torch.float64)
        done = torch.zeros(n, dtype=torch.uint8)
        if (t + 1) % T == 0:
            done[:] = 1
        wrote = log.step(act, obs, fin, done)
        assert wrote == (2 if done[0] else 0)
        hist.append((act, obs, fin, done))
    names = sorted(os.path.basename(p).split(".")[0] for p in log.paths)
    assert names == ["env000101_episode_1", "env000101_episode_2", "env000104_episode_1", "env000104_episode_2"]
    path = [p for p in log.paths if "env000104_episode_2" in p][0]
    from gym_kmanip_b200.log_episode import read_episode
    assert path.endswith(".hdf5")
    qpos, qvel, action, attrs, meta = read_episode(path)
    assert meta["env"] == 104 and meta["episode"] == 2 and meta["step"] == T and meta["env_id"] == "KManipSoloArm" and attrs["sim"]
    assert qpos.shape == (64, q_len) and action.shape == (64, 3) and qpos.dtype == np.float32
    for r in range(T):
        act, obs, fin, done = hist[T + r]
        src = fin if done[4] else obs
        assert np.allclose(qpos[r], src[4, :q_len].numpy(), atol=1e-7) and np.allclose(qvel[r], src[4, q_len:2 * q_len].numpy(), atol=1e-7)
        assert np.all(action[r] == act[4, 6].numpy())
    assert not qpos[T:].any() and not action[T:].any()
    assert int(log.row[0]) == 1 and log.qpos[1:].abs().sum() == 0   # the ninth step opened episode 3


def test_episode_log_with_cameras(tmp_path):
    """Vision ids: per-camera image datasets and intrinsics in the episode file (reference log_h5py.py:36-46, 59-60)."""
    from gym_kmanip_b200 import constants as K, log_episode
    cams = [K.CAMERAS["grip_r"], K.CAMERAS["grip_l"]]
    info = {"episode": 7, "sim": True, "q_len": 10, "a_len": 3, "step": 0, "cameras": cams}
    f = log_episode.new(str(tmp_path), info)
    rng = np.random.default_rng(2)
    frames = []
    for t in range(1, 4):
        info["step"] = t
        obs = {"q_pos": rng.uniform(0, 1, 10), "q_vel": rng.uniform(-1, 1, 10)}
        for c in cams:
            obs[c.log_name] = rng.integers(0, 256, (c.h, c.w, c.c), dtype=np.uint8)
        log_episode.step(f, {"grip_r": np.float32([0.5])}, obs, info)
        frames.append(obs)
    path = log_episode.end(f)
    qpos, qvel, action, attrs, meta, images, cam_meta = log_episode.read_episode(path, with_images=True)
    assert sorted(images) == ["grip_l", "grip_r"] and images["grip_r"].shape == (64, 40, 60, 3) and images["grip_r"].dtype == np.uint8
    for t, obs in enumerate(frames):
        assert np.array_equal(images["grip_r"][t], obs["camera/grip_r"]) and np.array_equal(images["grip_l"][t], obs["camera/grip_l"])
    assert not images["grip_r"][3:].any()
    assert list(cam_meta["camera/grip_r"]["resolution"]) == [60, 40] and cam_meta["camera/grip_r"]["focal_length"] == 45
    assert list(cam_meta["camera/grip_l"]["principal_point"]) == [30, 20] and "cameras" not in meta and meta["episode"] == 7
