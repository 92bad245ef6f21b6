"""CPU: host-side logic of the drop-in boundary -- registry, spaces, flat action / observation layouts, task structs,
env sharding -- mirroring what gymnasium's check_env pins for the reference (tests/test_env.py:8-24)."""
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


def test_vision_ids_and_loggers_are_declared_out_of_scope():
    from gym_kmanip_b200.env_base import KManipEnv
    with pytest.raises(NotImplementedError):
        KManipEnv(**K.ENV_REGISTRY["KManipSoloArmVision"])
    with pytest.raises(NotImplementedError):
        KManipEnv(**dict(K.ENV_REGISTRY["KManipSoloArm"], log_h5py=True))
    with pytest.raises(KeyError):
        k.make("KManipNope")
