"""CPU: the exact-parity IK mode (gym_kmanip_b200/csrc/km_ik_trf.cuh, the restated scipy TRF; host build of the kernel
source) against the REAL scipy.optimize.least_squares driving the oracle's ik_res / ik_jac exactly as the reference
does (ik_mujoco.py:129-135).  This is the one piece of the hot path whose third-party arithmetic is importable here,
so it is compared with the genuine article, not with a restatement."""
import numpy as np
import pytest

import hostsim
from oracle import oracle as om
from parity_util import rel_err

pytest.importorskip("scipy.optimize")


def _states(o, rng, n, spread):
    st0 = om.batch_reset_state(o, 1, seed=1)
    out = []
    rngs = np.array(o.flat["jnt_range"])[: o.nu]
    for _ in range(n):
        qpos = st0["qpos"][0].copy()
        qpos[: o.nu] = np.clip(qpos[: o.nu] + rng.uniform(-spread, spread, o.nu), rngs[:, 0] + 1e-4, rngs[:, 1] - 1e-4)
        out.append(qpos)
    return out, st0


@pytest.mark.parametrize("env_id", ["KManipSoloArm", "KManipDualArm", "KManipTorso"])
def test_before_step_with_restated_trf_matches_real_scipy(env_id):
    o = om.Oracle(env_id, ik_mode="trf")
    hs = hostsim.HostSim(env_id, 64, ik_mode=1)
    rng = np.random.default_rng(0)
    states, st0 = _states(o, rng, 10, 0.25)
    worst_ctrl = worst_qpos = 0.0
    nfev = []
    for qpos in states:
        if env_id == "KManipTorso":
            # the Torso home pose violates three joint limits: scipy raises for x0 out of bounds and the
            # reference then keeps the current joints; both sides must skip the solve for that arm
            pass
        st = dict(qpos=qpos, qvel=np.zeros(o.nv), ctrl=qpos[: o.nu].astype(np.float32).astype(np.float64), warm=np.zeros(o.nv),
                  mocap=st0["mocap"][0][: 7 * o.nmocap].copy(), time=0.0)
        act = rng.uniform(-1, 1, o.task.act_dim).astype(np.float32)
        o.set_state(st["qpos"], st["qvel"], st["ctrl"], st["warm"], 0.0, st["mocap"] if o.nmocap else None)
        o.ik_nfev.clear()
        o.before_step(act)
        nfev += o.ik_nfev
        ref = o.get_state()
        hs.set_state(st)
        hs.step1()
        hs.before_step(act)
        got = hs.get_state()
        worst_ctrl = max(worst_ctrl, float(np.abs(got["ctrl"] - ref["ctrl"]).max()))
        worst_qpos = max(worst_qpos, float(np.abs(got["qpos"] - ref["qpos"]).max()))
        assert np.allclose(got["mocap"], ref["mocap"], atol=1e-12)       # the IK goal written to the mocap body
    print(env_id, "worst |ctrl - scipy| %.2e  worst |qpos(teleported) - scipy| %.2e  scipy nfev %s" % (worst_ctrl, worst_qpos, sorted(set(nfev))))
    # ctrl is float32-rounded on both sides: agreement to one float32 ulp of a ~2 rad angle; qpos carries the fp64 solution
    assert worst_ctrl < 5e-7 and worst_qpos < 1e-7


def test_trf_env_step_matches_oracle_with_real_scipy():
    """Whole env step in exact-parity mode against the oracle whose IK is the real scipy TRF."""
    env_id = "KManipSoloArm"
    o = om.Oracle(env_id, ik_mode="trf")
    hs = hostsim.HostSim(env_id, 64, ik_mode=1)
    rng = np.random.default_rng(5)
    states, st0 = _states(o, rng, 4, 0.1)
    for qpos in states:
        st = dict(qpos=qpos, qvel=rng.normal(size=o.nv) * 0.1, ctrl=qpos[: o.nu].astype(np.float32).astype(np.float64),
                  warm=np.zeros(o.nv), mocap=st0["mocap"][0][: 7 * o.nmocap].copy(), time=0.0)
        act = rng.uniform(-1, 1, o.task.act_dim).astype(np.float32)
        o.set_state(st["qpos"], st["qvel"], st["ctrl"], st["warm"], 0.0, st["mocap"])
        obs, rew = o.step(act)
        hs.set_state(st)
        out = hs.env_step(act)
        a, b = hs.get_state(), o.get_state()
        assert rel_err(a["qpos"], b["qpos"]) < 1e-6 and rel_err(a["qvel"], b["qvel"], floor=1.0) < 1e-4
        assert rel_err(out["obs"], obs, floor=1.0) < 1e-4


def test_out_of_bounds_start_skips_the_solve_like_scipy_raises():
    """The Torso home pose violates joint limits: scipy raises ValueError for an infeasible x0, the
    reference swallows it and keeps the current joints (ik_mujoco.py:128-138); both IK modes must do the same."""
    env_id = "KManipTorso"
    o = om.Oracle(env_id, ik_mode="trf")
    st0 = om.batch_reset_state(o, 1, seed=3)
    qpos = st0["qpos"][0].copy()
    rngs = np.array(o.flat["jnt_range"])[: o.nu]
    mask_r = [o.task.arm_mask[0][i] for i in range(o.task.arm_nmask[0])]
    assert any(qpos[j] < rngs[j, 0] or qpos[j] > rngs[j, 1] for j in range(o.nu)), "home pose expected to violate a limit"
    st = dict(qpos=qpos, qvel=np.zeros(o.nv), ctrl=qpos[: o.nu].astype(np.float32).astype(np.float64), warm=np.zeros(o.nv),
              mocap=st0["mocap"][0][: 7 * o.nmocap].copy(), time=0.0)
    act = np.random.default_rng(1).uniform(-1, 1, o.task.act_dim).astype(np.float32)
    o.set_state(st["qpos"], st["qvel"], st["ctrl"], st["warm"], 0.0, st["mocap"])
    o.before_step(act)
    ref = o.get_state()
    for mode in (0, 1):
        hs = hostsim.HostSim(env_id, 64, ik_mode=mode)
        hs.set_state(st)
        hs.step1()
        hs.before_step(act)
        got = hs.get_state()
        infeasible = [a for a in range(o.task.n_arm)
                      if any(qpos[o.task.arm_mask[a][i]] < rngs[o.task.arm_mask[a][i], 0] or qpos[o.task.arm_mask[a][i]] > rngs[o.task.arm_mask[a][i], 1]
                             for i in range(o.task.arm_nmask[a]))]
        assert infeasible, "at least one arm starts out of bounds"
        for a in infeasible:   # that arm: qpos untouched, ctrl = clip(current joints) rounded to float32
            for i in range(o.task.arm_nmask[a]):
                j = o.task.arm_mask[a][i]
                assert got["qpos"][j] == qpos[j]
                assert got["ctrl"][j] == np.float32(np.clip(qpos[j], rngs[j, 0], rngs[j, 1]))
        if mode == 1:
            assert np.abs(got["ctrl"] - ref["ctrl"]).max() < 5e-7 and np.abs(got["qpos"] - ref["qpos"]).max() < 1e-7
