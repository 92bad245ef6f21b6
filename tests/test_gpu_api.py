"""GPU: the reference-facing Python API (KManipEnv / make, KManipVectorEnv) and size-independent properties of the
CUDA path at BASELINE.json's batch sizes.  The API checks restate what gymnasium's check_env pins for the reference
(tests/test_env.py:8-24: spaces, dtypes, bounds, 5-tuple types), plus the TimeLimit truncation of gym.make."""
import os

import numpy as np
import pytest

import gym_kmanip_b200 as k

STATE_IDS = ["KManipSoloArm", "KManipSoloArmQPos", "KManipDualArm", "KManipDualArmQPos", "KManipTorso"]


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", STATE_IDS)
def test_single_env_follows_the_reference_contract(env_id):
    env = k.make(env_id)
    u = env.unwrapped
    obs, info = env.reset(seed=0)
    assert list(obs) == ["q_pos", "q_vel", "cube_pos", "cube_orn"]
    assert u.observation_space.contains(obs), obs
    for key in ("step", "episode", "is_success", "q_keys", "q_len", "a_len", "obs_list", "act_list", "cameras", "sim",
                "sim_time", "cpu_time", "reward", "terminated"):
        assert key in info
    assert info["episode"] == 1 and info["step"] == 0 and info["reward"] is None and info["sim_time"] == 0.0
    home = k.ENV_REGISTRY[env_id]["q_pos_home"]
    qpos0 = u.env.physics.data.qpos
    assert np.allclose(qpos0[: u.q_len], home.astype(np.float64)) and np.allclose(qpos0[-4:], [1, 0, 0, 0])
    assert 0.1 <= qpos0[-7] <= 0.3 and 0.5 <= qpos0[-6] <= 0.7 and 0.6 <= qpos0[-5] <= 0.7      # CUBE_SPAWN_RANGE
    u.action_space.seed(0)
    for t in range(k.MAX_EPISODE_STEPS):
        a = u.action_space.sample()
        obs, reward, terminated, truncated, info = env.step(a)
        assert u.observation_space.contains(obs)
        assert all(v.dtype == np.float64 for v in obs.values())
        assert isinstance(reward, float) and np.isfinite(reward)
        assert terminated is False and isinstance(truncated, bool)
        assert truncated == (t == k.MAX_EPISODE_STEPS - 1)           # TimeLimit of gym.make (__init__.py:28,247)
        assert info["step"] == t + 1 and info["is_success"] == (reward > 2.0)
        assert abs(info["sim_time"] - 0.02 * (t + 1)) < 1e-9          # 10 sub-steps of 2 ms per env step
    d = u.env.physics.data
    assert d.qpos.shape == (u.env.physics.model.nq,) and d.qvel.shape == (u.env.physics.model.nv,)
    assert abs(np.linalg.norm(d.qpos[-4:]) - 1) < 1e-9
    site = d.site("eer_site_pos")
    assert site.xpos.shape == (3,) and site.xmat.shape == (9,)
    if "eer_pos" in u.act_list:   # the mocap body carries the last IK goal (env_sim.py:67-70)
        assert np.linalg.norm(d.mocap_pos[0] - site.xpos) < 0.1
    obs2, info2 = env.reset()
    assert info2["episode"] == 2 and info2["step"] == 0
    env.close()


@pytest.mark.gpu
def test_single_env_matches_oracle_free_running():
    """The single-env class end to end (dict actions in, dict observations out; fp64, exact-parity IK) against the oracle
    whose IK is the real scipy TRF, free-running from the same spawn (agreement stays far below contact chaos)."""
    pytest.importorskip("scipy.optimize")
    from oracle import oracle as om
    env = k.make("KManipSoloArm")
    o = om.Oracle("KManipSoloArm", ik_mode="trf")
    xyz = np.array([0.22, 0.61, 0.63])
    obs, _ = env.reset(options={"cube_xyz": xyz})
    assert np.allclose(np.concatenate(list(obs.values())), o.reset(xyz), atol=1e-12)
    rng = np.random.default_rng(1)
    for t in range(20):
        a = rng.uniform(-1, 1, 7).astype(np.float32)
        obs, rew, _, _, _ = env.step({"eer_pos": a[0:3], "eer_orn": a[3:6], "grip_r": a[6:7]})
        o_obs, o_rew = o.step(a)
        assert np.abs(np.concatenate(list(obs.values())) - o_obs).max() < 1e-7 and abs(rew - o_rew) < 1e-7
    env.close()


@pytest.mark.gpu
def test_vector_env_autoreset_and_totals():
    import torch
    n = 512
    env = k.make_vec("KManipDualArm", n, dtype="float32", seed=5)
    obs, _ = env.reset()
    assert list(obs) == ["q_pos", "q_vel", "cube_pos", "cube_orn"] and obs["q_pos"].shape == (n, 20) and obs["cube_orn"].shape == (n, 4)
    first = {kk: v.clone() for kk, v in obs.items()}
    # batched spaces, as gymnasium.vector.VectorEnv exposes them
    assert env.observation_space.spaces["q_pos"].shape == (n, 20) and env.action_space.spaces["eer_pos"].shape == (n, 3)
    assert env.single_action_space.spaces["grip_l"].shape == (1,) and env.observation_space.contains({kk: v.cpu().numpy() for kk, v in obs.items()})
    gen = torch.Generator(device="cuda").manual_seed(0)
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    tot = np.zeros(4)
    prev_obs = None
    for t in range(k.MAX_EPISODE_STEPS + 2):
        flat = env.sample_actions(gen)
        act = {kk: flat[:, sl] for kk, sl in env.action_layout.items()}       # dict actions with the reference keys
        obs, rew, term, trunc, info = env.step(act)
        if prev_obs is not None:     # default copy=True: what step() returned before is not overwritten by the next step
            assert torch.equal(prev_obs[0]["q_pos"], prev_obs[1])
        prev_obs = (obs, obs["q_pos"].clone())
        assert not term.any()
        assert bool(trunc.all()) == (t == k.MAX_EPISODE_STEPS - 1) and bool(trunc.any()) == (t == k.MAX_EPISODE_STEPS - 1)
        for v in obs.values():
            assert float(v.min()) >= -1 and float(v.max()) <= 1
        # the batched `info` of reference env_base.py:243-250, produced by the step kernel
        in_ep = t % k.MAX_EPISODE_STEPS + 1
        assert (info["step"] == in_ep).all() and (info["episode"] == t // k.MAX_EPISODE_STEPS).all()
        assert torch.allclose(info["sim_time"].double(), torch.full((n,), in_ep * k.CONTROL_TIMESTEP, dtype=torch.float64, device="cuda"), atol=1e-4)
        assert torch.equal(info["is_success"], rew > k.REWARD_SUCCESS_THRESHOLD)
        ret += rew.double()
        assert torch.allclose(info["episode_return"].double() * (~trunc) + info["final_return"].double() * trunc, ret, rtol=1e-4, atol=1e-4)
        tot += [float(rew.double().sum()), n, float(trunc.sum()), float(info["is_success"].sum())]
        if t == k.MAX_EPISODE_STEPS - 1:
            # same-step autoreset: obs is the first observation of the new episode, final_obs the last of the old one
            assert torch.equal(obs["q_pos"], first["q_pos"]) and torch.equal(obs["q_vel"], first["q_vel"])
            assert not torch.equal(obs["cube_pos"], first["cube_pos"])       # new spawn for episode 1
            assert not torch.equal(info["final_obs"]["q_pos"], first["q_pos"])
            assert (info["final_return"] != 0).all() and torch.equal(info["final_return"], info["episode_return"])
            ret.zero_()
        else:
            assert (info["final_return"] == 0).all()
    got = env.totals.cpu().numpy()           # km_episode_stats: accumulated by the step kernel, no extra launches
    assert got[1] == n * (k.MAX_EPISODE_STEPS + 2) and got[2] == n and got[3] == tot[3]
    assert abs(got[0] - tot[0]) < 1e-3 * max(1.0, abs(tot[0]))
    assert float(env.sim.episode_stats(reset=True)[1]) == got[1] and float(env.totals[1]) == 0
    env.close()


@pytest.mark.gpu
def test_vector_env_step_is_one_kernel_launch():
    """copy=False: step() returns views into the simulator's buffers and launches exactly ONE kernel -- the step kernel
    (km_launch_count, and the CUDA activity of the call as the torch profiler sees it, when CUPTI is available)."""
    import torch
    n = 256
    env = k.make_vec("KManipSoloArmQPos", n, dtype="float32", seed=1, copy=False)
    env.reset()
    act = env.sample_actions(torch.Generator(device="cuda").manual_seed(0))
    env.step(act)
    l0 = env.sim.launches
    obs, rew, term, trunc, info = env.step(act)
    assert env.sim.launches - l0 == 1
    assert obs["q_pos"].data_ptr() == env.sim.obs.data_ptr() and rew.data_ptr() == env.sim.reward.data_ptr()
    assert info["step"].data_ptr() == env.sim.step_count.data_ptr()
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            env.step(act)
            torch.cuda.synchronize()
        kernels = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    except Exception:
        kernels = None
    if kernels:      # CUPTI present: the only kernel of the call is ours
        assert len(kernels) == 1 and "k_env_step" in kernels[0], kernels
    env.close()


@pytest.mark.gpu
def test_batch_sim_argument_validation():
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    with pytest.raises(ValueError):
        BatchSim("KManipSoloArm", 8, dtype="fp32")
    sim = BatchSim("KManipSoloArm", 8, dtype="float32")
    sim.reset()
    with pytest.raises(ValueError):
        sim.reset(mask=torch.ones(7, dtype=torch.uint8))
    with pytest.raises(ValueError):
        sim.step(torch.zeros(8, sim.act_dim + 1))
    sim.step(torch.zeros(8, sim.act_dim))          # a CPU tensor is moved to the handle's device
    sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,n", [("KManipSoloArmQPos", 4096), ("KManipSoloArm", 8192)])
@pytest.mark.parametrize("lanes", [32, 2])
def test_properties_at_baseline_batch_sizes(env_id, n, lanes):
    """Size-independent properties at the BASELINE.json batch sizes: bit-exact determinism, independence from how envs
    are sharded into handles (global env ids; same kernel mapping on every shard -- km_create would pick the mapping
    from each shard's own size), unit quaternions, observation bounds, truncation every 64 steps."""
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    whole = BatchSim(env_id, n, dtype="float32", seed=9)
    again = BatchSim(env_id, n, dtype="float32", seed=9)
    half = [BatchSim(env_id, n // 2, dtype="float32", seed=9, env0=0), BatchSim(env_id, n // 2, dtype="float32", seed=9, env0=n // 2)]
    for s in [whole, again] + half:
        s.configure(lanes, 0)
        s.reset()
    gen = torch.Generator(device="cuda").manual_seed(3)
    ntrunc = 0
    for t in range(70):
        act = torch.rand(n, whole.act_dim, device="cuda", generator=gen) * 2 - 1
        obs, rew, term, trunc = whole.step(act, autoreset=True)
        obs2, rew2, _, _ = again.step(act, autoreset=True)
        assert torch.equal(obs, obs2) and torch.equal(rew, rew2)
        oh = torch.cat([half[0].step(act[: n // 2].contiguous(), autoreset=True)[0], half[1].step(act[n // 2:].contiguous(), autoreset=True)[0]])
        assert torch.equal(obs, oh)
        ntrunc += int(trunc.sum())
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        assert float(obs.min()) >= -1 and float(obs.max()) <= 1
        quat = obs[:, -4:]
        assert float((quat.norm(dim=1) - 1).abs().max()) < 1e-5
    assert ntrunc == n
    for s in [whole, again] + half:
        s.close()


@pytest.mark.gpu
def test_mappings_agree_with_each_other():
    """The three kernel mappings (lane group, thread per env in shared / local memory) are the same simulator: from
    the same state and action they agree to fp32 rounding."""
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    n = 512
    sims = []
    for lanes in (32, 1, 2):
        s = BatchSim("KManipSoloArm", n, dtype="float32", seed=4)
        s.configure(lanes, 0)
        s.reset()
        sims.append(s)
    gen = torch.Generator(device="cuda").manual_seed(1)
    for t in range(6):
        act = torch.rand(n, sims[0].act_dim, device="cuda", generator=gen) * 2 - 1
        outs = [s.step(act, autoreset=True)[0].clone() for s in sims]
        st0 = sims[0].get_state()[0]
        for s, o in zip(sims[1:], outs[1:]):
            assert float((o - outs[0]).abs().max()) < 2e-3          # q_vel / pi entries carry ~100 rad/s velocities
            s.set_state(st0)                                          # teacher-force: keep the comparison per step
    for s in sims:
        s.close()


@pytest.mark.gpu
def test_single_env_episode_logging(tmp_path, monkeypatch):
    """log_h5py=True (reference examples/2_log_with_h5py.py): one file per episode under log_dir, reference layout."""
    from gym_kmanip_b200 import constants as K
    monkeypatch.setattr(K, "DATA_DIR", str(tmp_path))
    env = k.make("KManipSoloArm", log_h5py=True, ik_mode=0)
    u = env.unwrapped
    u.action_space.seed(0)
    env.reset()
    for _ in range(5):
        env.step(u.action_space.sample())
    env.reset()
    env.step(u.action_space.sample())
    env.close()
    files = sorted(os.listdir(u.log_dir))
    assert [f.split(".")[0] for f in files] == ["episode_1", "episode_2"]
    from gym_kmanip_b200.log_episode import read_episode
    assert files[0].endswith(".hdf5")
    qpos, qvel, action, attrs, meta = read_episode(os.path.join(u.log_dir, files[0]))
    assert qpos.shape == (64, 10) and action.shape == (64, 3) and attrs["sim"]
    assert qpos[:5].any() and not qpos[5:].any()


@pytest.mark.gpu
def test_single_env_episode_logging_with_cameras(tmp_path, monkeypatch):
    """Vision id with log_h5py=True: the episode file carries the camera observations and intrinsics
    (reference log_h5py.py:36-46, 59-60; env_base.py:231-234)."""
    from gym_kmanip_b200 import constants as K
    from gym_kmanip_b200.log_episode import read_episode
    monkeypatch.setattr(K, "DATA_DIR", str(tmp_path))
    env = k.make("KManipSoloArmVision", log_h5py=True, ik_mode=0)
    u = env.unwrapped
    u.action_space.seed(0)
    obs0, _ = env.reset()
    seen = []
    for _ in range(3):
        obs, *_ = env.step(u.action_space.sample())
        seen.append(obs)
    env.close()
    files = sorted(os.listdir(u.log_dir))
    assert files == ["episode_1.hdf5"]
    qpos, qvel, action, attrs, meta, images, cam_meta = read_episode(os.path.join(u.log_dir, files[0]), with_images=True)
    assert sorted(images) == ["grip_r", "head"] and images["head"].shape == (64, 480, 640, 3) and images["grip_r"].shape == (64, 40, 60, 3)
    for t, obs in enumerate(seen):
        assert np.array_equal(images["head"][t], obs["camera/head"]) and np.array_equal(images["grip_r"][t], obs["camera/grip_r"])
    assert images["head"][:3].any() and not images["head"][3:].any()
    assert list(cam_meta["camera/head"]["resolution"]) == [640, 480] and cam_meta["camera/head"]["focal_length"] == 448


@pytest.mark.gpu
def test_vector_env_episode_logging_from_device_ring_buffers(tmp_path):
    """KManipVectorEnv(log_dir=...): logged envs' rows stay on the device until truncation, then one file per episode."""
    import torch
    from gym_kmanip_b200.vector_env import KManipVectorEnv
    env = KManipVectorEnv("KManipSoloArmQPos", 64, seed=3, max_episode_steps=4, log_dir=str(tmp_path), log_env_ids=[0, 63])
    obs, _ = env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []
    for t in range(9):
        act = env.sample_actions(g)
        obs, rew, term, trunc, info = env.step(act)
        src = torch.where(trunc.bool()[:, None], env.sim.final_obs, env.sim.obs)
        rows.append((act[63, env.action_layout["grip_r"].start].item(), src[63, :10].float().cpu().numpy()))
    files = sorted(os.listdir(tmp_path))
    assert [f.split(".")[0] for f in files] == ["env000000_episode_1", "env000000_episode_2", "env000063_episode_1", "env000063_episode_2"]
    from gym_kmanip_b200.log_episode import read_episode
    assert files[-1].endswith(".hdf5")
    qpos, qvel, action, attrs, meta = read_episode(os.path.join(tmp_path, files[-1]))
    assert qpos.shape == (64, 10) and action.shape == (64, 2)
    for r in range(4):
        assert np.allclose(qpos[r], rows[4 + r][1], atol=1e-7) and np.all(action[r] == np.float32(rows[4 + r][0]))
    assert not qpos[4:].any()
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,lanes,epb", [("KManipSoloArmQPos", 32, 2), ("KManipSoloArm", 32, 3), ("KManipSoloArmQPos", 2, 64)])
def test_cost_ordered_walk_changes_no_result(env_id, lanes, epb):
    """km_set_env_ordering: the envs are walked in the order of their previous step's solver cost and the CTAs fetch tiles
    dynamically; every env must come out bit-identical to the plain walk, the permutation must be a permutation (every env
    stepped exactly once per step), and the automatic mode must switch it on only when there are more tiles than CTAs."""
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    n = 1500
    sims = []
    for mode in (0, 1):
        sim = BatchSim(env_id, n, dtype="float32", seed=3)
        sim.configure(lanes, epb)
        sim.set_env_ordering(mode)
        sim.reset()
        sims.append(sim)
    gen = torch.Generator(device="cuda").manual_seed(7)
    for t in range(6):
        act = torch.rand(n, sims[0].act_dim, device="cuda", generator=gen) * 2 - 1
        outs = [[x.clone() for x in sim.step(act)] for sim in sims]
        for a, b in zip(*outs):
            assert torch.equal(a, b)
        assert torch.equal(sims[0].step_count, sims[1].step_count) and bool((sims[1].step_count == t + 1).all())
    assert torch.equal(sims[0].get_state()[0], sims[1].get_state()[0])
    l0 = [s.launches for s in sims]
    for s in sims:
        s.step(act)
    assert sims[0].launches - l0[0] == 1 and sims[1].launches - l0[1] == 2   # the counting sort is the second launch
    tot = [s.episode_stats().cpu() for s in sims]
    assert tot[0][1] == tot[1][1] == n * 7 and abs(float(tot[0][0] - tot[1][0])) < 1e-6 * abs(float(tot[0][0]))
    # automatic mode: 1500 envs in tiles of `epb` are more tiles than CTAs only for the small lane-group tiles
    auto = BatchSim(env_id, n, dtype="float32", seed=3)
    cfg = auto.configure(lanes, epb)
    auto.reset()
    l1 = auto.launches
    auto.step(act)
    tiles = (n + cfg["envs_per_block"] - 1) // cfg["envs_per_block"]
    assert auto.launches - l1 == (2 if (lanes >= 16 and tiles > cfg["grid"]) else 1)
    for s in sims + [auto]:
        s.close()
