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

from parity_util import (CONTACT_TOL_POS_F32, CONTACT_TOL_VEL_F32, comp_rel_err_per_env, component_floors, oracle_rollout,
                         pack_state, rel_err, rel_err_per_env)

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


# ------------------------------------------------------------------------------------------------ per sub-step, per component
def _quant(x):
    x = np.concatenate(x) if len(x) else np.zeros(0)
    if x.size == 0:
        return dict(p50=0.0, p99=0.0, max=0.0, n=0)
    return dict(p50=float(np.quantile(x, 0.5)), p99=float(np.quantile(x, 0.99)), max=float(x.max()), n=int(x.size))


def _sub_step_errors(env_id, dtype, n=256, steps=40):
    """Teacher-forced SINGLE physics sub-steps (km_task.n_sub_steps = 1: before_step, mj_step2, mj_step1) against the
    oracle running the same: distribution over env-sub-steps of the per-component relative error
    (parity_util.comp_rel_err_per_env: every component on the scale of its own unit) of the state after one mj_step."""
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    f32 = dtype == "float32"
    o, traj = oracle_rollout(env_id, n, steps, seed=7, action_seed=8, round32=f32, n_sub_steps=1)
    sim = BatchSim(env_id, n, dtype=dtype, seed=7, n_sub_steps=1)
    fq, fv = component_floors(o.flat, o.nq, o.nv)
    acc = dict(free_pos=[], free_vel=[], con_pos=[], con_vel=[])
    flips = touching = 0
    for rec in traj:
        b, a = rec["before"], rec["after"]
        sim.set_state(pack_state(b), step=b["step"], episode=b["episode"])
        sim.step(torch.from_numpy(rec["action"]).cuda(), autoreset=True)
        torch.cuda.synchronize()
        g = _sim_state(sim)
        touch = rec["ncon_peak"] > 0
        free = ~touch
        touching += int(touch.sum())
        ep, ev = comp_rel_err_per_env(g["qpos"], a["qpos"], fq), comp_rel_err_per_env(g["qvel"], a["qvel"], fv)
        acc["free_pos"].append(ep[free]); acc["free_vel"].append(ev[free])
        acc["con_pos"].append(ep[touch]); acc["con_vel"].append(ev[touch])
        mc = sim.max_contacts
        same = (sim.ncon.cpu().numpy() == rec["ncon"]) & (sim.con_geoms.cpu().numpy() == rec["geoms"][:, : 2 * mc]).all(axis=1)
        assert same[free].all(), "contact pairs of an env whose cube touches nothing must be bit-exact"
        flips += int((~same).sum())
        assert np.array_equal(g["step"], a["step"]) and np.array_equal(sim.truncated.cpu().numpy(), rec["truncated"])
    sim.close()
    w = {k: _quant(v) for k, v in acc.items()}
    w["flips"], w["touching"] = flips, touching
    return w


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["KManipSoloArmQPos", "KManipSoloArm", "KManipDualArm", "KManipTorso"])
def test_single_sub_step_parity_fp32_per_component(env_id):
    """BASELINE.json north_star: "per-step qpos/qvel within 1e-5 relative in the fp32 build ... bit-exact contact-pair
    indices".  One physics sub-step (mj_step) from identical float32-representable states, every component on the scale
    of its own unit, ~10 000 env-sub-steps per scene.  Envs whose cube touches nothing:
      positions  -- 99 % of the env-sub-steps within 1e-5 on every component (median ~1e-7), none above 1e-4;
      velocities -- get h * qacc, and qacc comes out of an fp32 factorisation of the servo-stiff Newton system (kp = 1000
                    on 50 g wrist links, limit rows with D ~ 1e6: |qacc| ~ 1e3..1e4 rad/s^2 with relative error
                    cond(H) * eps): 1e-5 on velocities is not attainable in float32.  99 % within 2e-4 of the component's
                    own scale, none above 5e-3; the fp64 build holds 1e-10 (next test).
    Contact pairs bit-exact for those envs.  Envs in contact: the cube rests ~1e-7 m deep in the table, below float32
    resolution of its height (DESIGN.md "fp32 and the cube"): looser bounds, contact-report flips must stay rare.
    The achieved figures are printed; profiles/r02_notes.md records them (also for a build without
    -prec-div=false -prec-sqrt=false -ftz=true)."""
    w = _sub_step_errors(env_id, "float32")
    print(env_id, "fp32 one sub-step:", {k: ({kk: (f"{vv:.1e}" if isinstance(vv, float) else vv) for kk, vv in v.items()} if isinstance(v, dict) else v) for k, v in w.items()})
    assert w["free_pos"]["n"] > 1000
    assert w["free_pos"]["p99"] < 1e-5 and w["free_pos"]["max"] < 1e-4, w
    assert w["free_vel"]["p99"] < 2e-4 and w["free_vel"]["max"] < 5e-3, w
    assert w["con_pos"]["max"] < 1e-3 and w["con_vel"]["max"] < 0.2, w
    assert w["flips"] <= max(2, w["touching"] // 100), w


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["KManipSoloArm", "KManipTorso"])
def test_single_sub_step_parity_fp64_per_component(env_id):
    w = _sub_step_errors(env_id, "float64", n=128, steps=30)
    assert w["free_pos"]["max"] < 1e-12 and w["free_vel"]["max"] < 1e-10 and w["con_pos"]["max"] < 1e-12 and w["con_vel"]["max"] < 1e-9 and w["flips"] == 0, w


# ------------------------------------------------------------------------------------------------ production instantiations
def _tiled_case(env_id, n, dtype, t_mid, seed=13):
    """Teacher-forced inputs at a BASELINE batch size: the states of a small oracle rollout (fresh and mid-episode, cube
    resting on the table) tiled to n envs, with fresh random actions for every env; the oracle steps all n envs once."""
    from oracle import oracle as om
    f32 = dtype == "float32"
    n0 = 512
    o, traj = oracle_rollout(env_id, n0, t_mid + 1, seed=seed, action_seed=seed + 1, round32=f32)
    cases = []
    for t in (1, t_mid):
        b = traj[t]["before"]
        reps = (n + n0 - 1) // n0
        st = {k: np.ascontiguousarray(np.concatenate([v] * reps, axis=0)[:n]) for k, v in b.items()}
        act = np.random.default_rng(seed + 2 + t).uniform(-1, 1, (n, o.task.act_dim)).astype(np.float32)
        before = {k: v.copy() for k, v in st.items()}
        peak = np.zeros(n, dtype=np.int32)
        obs, fobs, rew, trunc, flags, ncon, geoms = om.batch_step(o, st, act, autoreset=True, seed=seed, nthreads=0, ncon_peak=peak)
        cases.append(dict(before=before, action=act, after=st, obs=obs, reward=rew, truncated=trunc, ncon=ncon, geoms=geoms, flags=flags, ncon_peak=peak))
    return o, cases


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,n,dtype,lanes", [
    ("KManipSoloArmQPos", 4096, "float32", 0),    # BASELINE configs[1]: default mapping (warp per env, register-resident solver)
    ("KManipSoloArm", 65536, "float32", 0),       # configs[2] total on one GPU: thread per env, 128-register 512-thread instantiation
    ("KManipDualArm", 32768, "float32", 0),       # configs[3]: thread per env, 222-thread CTAs
    ("KManipTorso", 16384, "float64", 0),         # configs[4]: fp64 validation build
    ("KManipSoloArmQPos", 4096, "float32", 16),   # two envs per warp
    ("KManipSoloArm", 8192, "float64", 0),        # fp64 at a production batch
])
def test_parity_at_baseline_batch_sizes_default_mapping(env_id, n, dtype, lanes):
    """The instantiations that actually run the BASELINE.json configurations (the mapping km_create picks at that batch
    size) against the oracle, teacher-forced: a fresh step and a mid-episode step with the cube resting on the table."""
    import torch
    from gym_kmanip_b200.batch_sim import BatchSim
    f32 = dtype == "float32"
    o, cases = _tiled_case(env_id, n, dtype, t_mid=22)
    sim = BatchSim(env_id, n, dtype=dtype, seed=13)
    if lanes:
        sim.configure(lanes, 0)
    cfg = sim.launch_config()
    for rec in cases:
        b, a = rec["before"], rec["after"]
        sim.set_state(pack_state(b), step=b["step"], episode=b["episode"])
        obs, rew, term, trunc = sim.step(torch.from_numpy(rec["action"]).cuda(), autoreset=True)
        torch.cuda.synchronize()
        g = _sim_state(sim)
        touch = (rec["ncon_peak"] > 0) if f32 else np.zeros(n, dtype=bool)
        free = ~touch
        tol_p, tol_v, tol_o = (2e-5, 1e-4, 1e-3) if f32 else (1e-10, 1e-10, 1e-9)
        # 10^4..10^5 envs per step here against 64 in the small tests: the fp32 bounds are asserted on the 99.9 % quantile
        # of the per-env errors, and the worst env must stay within 10x of them (fp64: every env within the bound)
        ep, ev = rel_err_per_env(g["qpos"][free], a["qpos"][free]), rel_err_per_env(g["qvel"][free], a["qvel"][free])
        qp, qv = (np.quantile(ep, 0.999), np.quantile(ev, 0.999)) if f32 else (ep.max(), ev.max())
        print(env_id, n, dtype, "qpos q99.9 %.1e max %.1e | qvel q99.9 %.1e max %.1e" % (np.quantile(ep, 0.999), ep.max(), np.quantile(ev, 0.999), ev.max()))
        assert qp < tol_p and qv < tol_v and ep.max() < 10 * tol_p and ev.max() < 10 * tol_v, cfg
        eo = rel_err_per_env(obs.double().cpu().numpy()[free], rec["obs"][free])
        assert (np.quantile(eo, 0.999) if f32 else eo.max()) < tol_o and eo.max() < 10 * tol_o
        assert rel_err(rew.double().cpu().numpy()[free], rec["reward"][free]) < 10 * tol_o
        if touch.any():
            assert np.abs(g["qpos"][touch] - a["qpos"][touch]).max() < CONTACT_TOL_POS_F32
            assert np.abs(g["qvel"][touch] - a["qvel"][touch]).max() < CONTACT_TOL_VEL_F32
        mc = sim.max_contacts
        same = (sim.ncon.cpu().numpy() == rec["ncon"]) & (sim.con_geoms.cpu().numpy() == rec["geoms"][:, : 2 * mc]).all(axis=1)
        assert same[free].all()
        assert (~same).sum() <= max(2, int(touch.sum()) // 100)
        assert np.array_equal(trunc.cpu().numpy(), rec["truncated"])
    print(env_id, n, dtype, "mapping", cfg)
    sim.close()
