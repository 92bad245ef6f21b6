"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle from identical states (teacher-forced).

Tolerances (parity_util.rel_err: max |a - b| over a batch field / that field's own magnitude over the batch):
  fp64 build: 1e-10 per env step on the state (BASELINE.json north_star); 1e-9 on the observation record, whose own
              scale is 1 while its q_vel entries carry the absolute error of velocities of magnitude ~30 rad/s;
              contact-pair indices, contact counts and done flags bit-exact for every env.
  fp32 build: teacher states are float32-representable (oracle_rollout(round32=True)), so no input-rounding error
              enters.  Envs whose cube touches nothing: 2e-5 on positions / 1e-4 on velocities (1e-3 absolute on the observation record) relative to the batch
              magnitude (~100 rad/s) -- the north_star 1e-5 figure is per physics sub-step, an env step chains ten.
              Envs in contact: absolute bounds CONTACT_TOL_*_F32 (parity_util; DESIGN.md "fp32 and the cube"), and
              bit-exact contact indices except where the oracle itself changes its answer under a one-ulp change of
              the cube height (those env-steps are counted and must stay rare).
"""
import os

import numpy as np
import pytest

from parity_util import CONTACT_TOL_POS_F32, CONTACT_TOL_VEL_F32, oracle_rollout, pack_state, rel_err

ENVS = ["KManipSoloArmQPos", "KManipSoloArm", "KManipDualArm", "KManipDualArmQPos", "KManipTorso"]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sim_state(sim):
    """state record -> dict of float64 arrays with the cube position recombined (hi + lo)."""
    st, stepc, ep = sim.get_state()
    st = st.double().cpu().numpy()
    sl = sim.state_slices()
    out = {k: st[:, s].copy() for k, s in sl.items()}
    out["qpos"][:, -7:-4] += out["cube_lo"]
    out["step"], out["episode"] = stepc.cpu().numpy(), ep.cpu().numpy()
    return out


def _run(env_id, dtype, n, steps, tol_pos, tol_vel, tol_obs, lanes=0):
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    f32 = dtype == "float32"
    o, traj = oracle_rollout(env_id, n, steps, seed=3, action_seed=5, round32=f32)
    sim = BatchSim(env_id, n, dtype=dtype, seed=3)
    if lanes:
        assert sim.configure(lanes, 0)["lanes_per_env"] == lanes
    worst = dict(qpos=0.0, qvel=0.0, ctrl=0.0, obs=0.0, reward=0.0, c_pos=0.0, c_vel=0.0)
    n_touch = n_flip = 0
    for t, rec in enumerate(traj):
        b = rec["before"]
        sim.set_state(pack_state(b), step=b["step"], episode=b["episode"])
        act = torch.from_numpy(rec["action"]).cuda()
        obs, rew, term, trunc = sim.step(act, autoreset=True)
        torch.cuda.synchronize()
        g = _sim_state(sim)
        a = rec["after"]
        obs_n, rew_n = obs.double().cpu().numpy(), rew.double().cpu().numpy()
        touch = (rec["ncon_peak"] > 0) if f32 else np.zeros(n, dtype=bool)
        free = ~touch
        n_touch += int(touch.sum())
        if free.any():
            errs = dict(qpos=rel_err(g["qpos"][free], a["qpos"][free]), qvel=rel_err(g["qvel"][free], a["qvel"][free]),
                        ctrl=rel_err(g["ctrl"][free], a["ctrl"][free]), obs=rel_err(obs_n[free], rec["obs"][free]),
                        reward=rel_err(rew_n[free], rec["reward"][free]))
            for k, v in errs.items():
                worst[k] = max(worst[k], v)
        if touch.any():
            worst["c_pos"] = max(worst["c_pos"], float(np.abs(g["qpos"][touch] - a["qpos"][touch]).max()))
            worst["c_vel"] = max(worst["c_vel"], float(np.abs(g["qvel"][touch] - a["qvel"][touch]).max()))
        # bit-exact integer outputs
        assert np.array_equal(trunc.cpu().numpy(), rec["truncated"]), f"truncated differs at step {t}"
        assert not term.any()
        assert np.array_equal(g["step"], a["step"]) and np.array_equal(g["episode"], a["episode"])
        ncon_g, geoms_g, flags_g = sim.ncon.cpu().numpy(), sim.con_geoms.cpu().numpy(), sim.con_flags.cpu().numpy()
        mc = sim.max_contacts
        same = (ncon_g == rec["ncon"]) & (geoms_g == rec["geoms"][:, : 2 * mc]).all(axis=1) & (flags_g == rec["flags"])
        if f32:
            n_flip += int((~same).sum())
            assert same[free].all(), f"contact report of a free env differs at step {t}"
        else:
            assert same.all(), f"contact report differs at step {t}"
    print(env_id, dtype, {k: "%.2e" % v for k, v in worst.items()}, f"touching env-steps {n_touch}, contact-report flips {n_flip}")
    assert worst["qpos"] < tol_pos and worst["ctrl"] < tol_pos, worst
    assert worst["qvel"] < tol_vel, worst
    assert worst["obs"] < tol_obs and worst["reward"] < tol_obs, worst
    assert worst["c_pos"] < CONTACT_TOL_POS_F32 and worst["c_vel"] < CONTACT_TOL_VEL_F32, worst
    assert n_flip <= max(2, n_touch // 100), (n_flip, n_touch)
    assert sim.launches >= steps
    sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ENVS)
def test_env_step_parity_fp64(env_id):
    _run(env_id, "float64", n=64, steps=70, tol_pos=1e-10, tol_vel=1e-10, tol_obs=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ENVS)
def test_env_step_parity_fp32(env_id):
    _run(env_id, "float32", n=64, steps=70, tol_pos=2e-5, tol_vel=1e-4, tol_obs=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["KManipSoloArm", "KManipDualArm", "KManipTorso", "KManipSoloArmQPos"])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("lanes", [1, 2])
def test_thread_per_env_mapping_parity(env_id, dtype, lanes):
    """The thread-per-env mapping (k_env_step_tpe with its own Newton solver; km_configure lanes 1: env records in
    shared memory, lanes 2: in local memory -- the default for large batches) against the oracle, same tolerances as
    the lane-group mapping."""
    if dtype == "float64":
        _run(env_id, dtype, n=64, steps=70, tol_pos=1e-10, tol_vel=1e-10, tol_obs=1e-9, lanes=lanes)
    else:
        _run(env_id, dtype, n=64, steps=70, tol_pos=2e-5, tol_vel=1e-4, tol_obs=1e-3, lanes=lanes)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ENVS)
def test_against_committed_golden_records_fp64(env_id):
    """The CUDA path against the committed fixtures (tests/golden/traj_*.npz, generated by make_golden.py from the
    oracle): no oracle code runs in this test."""
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    g = np.load(os.path.join(GOLD, f"traj_{env_id}.npz"))
    n = g["s0_before"].shape[0]
    sim = BatchSim(env_id, n, dtype="float64", seed=11)
    for t in g["steps"]:
        sim.set_state(g[f"s{t}_before"], step=g[f"s{t}_before_step"], episode=g[f"s{t}_before_episode"])
        obs, rew, term, trunc = sim.step(torch.from_numpy(g[f"s{t}_action"]).cuda(), autoreset=True)
        torch.cuda.synchronize()
        st, stepc, ep = sim.get_state()
        assert rel_err(st.cpu().numpy(), g[f"s{t}_after"]) < 1e-10
        assert rel_err(obs.cpu().numpy(), g[f"s{t}_obs"]) < 1e-9 and rel_err(rew.cpu().numpy(), g[f"s{t}_reward"]) < 1e-9
        tr = g[f"s{t}_truncated"]
        assert np.array_equal(trunc.cpu().numpy(), tr)
        if tr.any():
            assert rel_err(sim.final_obs.cpu().numpy()[tr != 0], g[f"s{t}_final_obs"][tr != 0]) < 1e-9
        assert np.array_equal(sim.ncon.cpu().numpy(), g[f"s{t}_ncon"])
        assert np.array_equal(sim.con_geoms.cpu().numpy(), g[f"s{t}_geoms"][:, : 2 * sim.max_contacts])
        assert np.array_equal(sim.con_flags.cpu().numpy(), g[f"s{t}_flags"])
        assert np.array_equal(stepc.cpu().numpy(), g[f"s{t}_after_step"]) and np.array_equal(ep.cpu().numpy(), g[f"s{t}_after_episode"])
    sim.close()


@pytest.mark.gpu
def test_reset_and_site_poses_match_oracle():
    """km_reset (initialize_episode, env_sim.py:23-36) with the device generator, and km_site_poses."""
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    from oracle import oracle as om
    for env_id in ("KManipSoloArm", "KManipDualArm", "KManipTorso"):
        n = 33
        o = om.Oracle(env_id)
        sim = BatchSim(env_id, n, dtype="float64", seed=17, env0=1000)
        obs = sim.reset().cpu().numpy()
        st = om.batch_reset_state(o, n, seed=17, env0=1000)
        g = _sim_state(sim)
        assert rel_err(g["qpos"], st["qpos"]) < 1e-15 and rel_err(g["ctrl"], st["ctrl"]) < 1e-15
        assert (g["qvel"] == 0).all() and (g["step"] == 0).all() and (g["episode"] == 0).all()
        pos, mat = sim.site_poses()
        pos, mat = pos.cpu().numpy(), mat.cpu().numpy()
        for i in range(0, n, 8):
            o.set_state(st["qpos"][i], st["qvel"][i], st["ctrl"][i])
            assert rel_err(obs[i], o.obs()) < 1e-12
            for a in range(o.task.n_arm):
                sid = o.task.arm_site[a]
                assert rel_err(pos[i, a], o.field("site_xpos").reshape(-1, 3)[sid]) < 1e-12
                assert rel_err(mat[i, a].reshape(9), o.field("site_xmat").reshape(-1, 9)[sid]) < 1e-12
        # masked reset: only the selected envs start a new episode (and get a new spawn)
        mask = torch.zeros(n, dtype=torch.uint8)
        mask[::3] = 1
        sim.reset(mask=mask.cuda())
        g2 = _sim_state(sim)
        assert (g2["episode"] == mask.numpy()).all()
        keep = mask.numpy() == 0
        assert np.array_equal(g2["qpos"][keep], g["qpos"][keep]) and not np.allclose(g2["qpos"][~keep, -7:-4], g["qpos"][~keep, -7:-4])
        sim.close()


@pytest.mark.gpu
def test_host_buffer_entry_points_equal_device_entry_points():
    """km_step_host / km_reset_host (what a binding that owns no device memory calls) == km_step / km_reset."""
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    n = 257
    a = BatchSim("KManipSoloArm", n, dtype="float32", seed=4)
    b = BatchSim("KManipSoloArm", n, dtype="float32", seed=4)
    h_obs = np.zeros((n, a.obs_dim), dtype=np.float32)
    h_rew, h_tr = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.uint8)
    a.reset()
    b.reset_host(None, None, h_obs)
    assert np.array_equal(a.obs.cpu().numpy(), h_obs)
    rng = np.random.default_rng(0)
    for t in range(66):
        act = rng.uniform(-1, 1, (n, a.act_dim)).astype(np.float32)
        obs, rew, term, trunc = a.step(torch.from_numpy(act).cuda(), autoreset=True)
        b.step_host(act, h_obs, h_rew, h_tr, autoreset=True)
        assert np.array_equal(obs.cpu().numpy(), h_obs) and np.array_equal(rew.cpu().numpy(), h_rew)
        assert np.array_equal(trunc.cpu().numpy(), h_tr)
    assert h_tr.sum() == 0 and a.launches == b.launches
    a.close()
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["KManipSoloArm", "KManipDualArm"])
def test_exact_trf_ik_mode_matches_real_scipy(env_id):
    """ik_mode = 1 (km_ik_trf.cuh on the device) against the oracle whose IK is the REAL scipy.optimize.least_squares,
    driven per env exactly as the reference does (ik_mujoco.py:129-135): whole env steps, fp64."""
    pytest.importorskip("scipy.optimize")
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    from oracle import oracle as om
    n, steps = 24, 3
    o = om.Oracle(env_id, ik_mode="trf")
    st = om.batch_reset_state(o, n, seed=2)
    if o.nmocap == 0:
        st["mocap"] = np.zeros((n, 0))
    sim = BatchSim(env_id, n, dtype="float64", seed=2, ik_mode=1)
    rng = np.random.default_rng(8)
    for t in range(steps):
        act = rng.uniform(-1, 1, (n, o.task.act_dim)).astype(np.float32)
        sim.set_state(pack_state(st), step=st["step"], episode=st["episode"])
        obs, rew, term, trunc = sim.step(torch.from_numpy(act).cuda(), autoreset=False)
        torch.cuda.synchronize()
        g = _sim_state(sim)
        obs_n = obs.cpu().numpy()
        for i in range(n):
            o.set_state(st["qpos"][i], st["qvel"][i], st["ctrl"][i], st["warm"][i], st["time"][i], st["mocap"][i] if o.nmocap else None)
            o_obs, o_rew = o.step(act[i])
            s = o.get_state()
            assert rel_err(g["qpos"][i], s["qpos"]) < 1e-9 and rel_err(g["qvel"][i], s["qvel"], floor=1.0) < 1e-8
            assert np.array_equal(g["ctrl"][i], s["ctrl"])            # float32-rounded on both sides: identical
            assert rel_err(obs_n[i], o_obs, floor=1.0) < 1e-8
            for k in ("qpos", "qvel", "ctrl", "warm"):
                st[k][i] = s[k]
            st["time"][i] = s["time"]
            if o.nmocap:
                st["mocap"][i] = s["mocap"]
        st["step"] += 1
    sim.close()
