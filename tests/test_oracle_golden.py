"""CPU: the oracle against the committed golden fixtures of tests/golden/ (see make_golden.py for what pins what)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as om
from parity_util import pack_state, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return json.load(open(os.path.join(GOLD, name)))


def test_euler_quaternion_conventions_match_scipy_fixture():
    """env_sim.py:62-66: as_euler("xyz") / from_euler("xyz").as_quat()[[3,0,1,2]] -- extrinsic xyz, wxyz order."""
    import ctypes as C
    g = _load("scipy_rotation.json")
    L = om.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))   # noqa: E731
    for mat, eul, quat in zip(g["mat"], g["euler_xyz"], g["quat_wxyz"]):
        m = np.ascontiguousarray(np.array(mat).reshape(9))
        e, q = np.zeros(3), np.zeros(4)
        L.ko_euler_pieces(dp(m), dp(e), dp(q))
        assert np.allclose(e, eul, atol=1e-12)
        q_ref = np.array(quat)
        assert min(np.abs(q - q_ref).max(), np.abs(q + q_ref).max()) < 1e-12


def test_subquat_matches_scipy_fixture():
    """ik_mujoco.py:43-46: mju_subQuat(goal, current) is the rotation vector of qb^-1 * qa."""
    import ctypes as C
    g = _load("scipy_rotation.json")
    L = om.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))   # noqa: E731
    eye = np.eye(3).reshape(9).copy()
    for qa, qb, sub in zip(g["qa_wxyz"], g["qb_wxyz"], g["subquat"]):
        a, b = np.array(qa), np.array(qb)
        qm, s = np.zeros(4), np.zeros(3)
        L.ko_quat_pieces(dp(eye), dp(a), dp(b), dp(qm), dp(s))
        assert np.allclose(s, sub, atol=1e-11)


def test_scipy_is_still_the_scipy_of_the_fixture():
    """The fixture was generated with the scipy installed in the image; re-derive one slice live when scipy imports."""
    R = pytest.importorskip("scipy.spatial.transform").Rotation
    g = _load("scipy_rotation.json")
    eul = R.from_matrix(np.array(g["mat"])).as_euler("xyz")
    assert np.allclose(eul, g["euler_xyz"], atol=1e-12)


@pytest.mark.parametrize("env_id", ["KManipSoloArm", "KManipDualArm", "KManipTorso"])
def test_fk_home_pose_matches_survey_numbers(env_id):
    """Site positions at the home pose vs the numbers derived independently during the survey."""
    g = _load("fk_home.json")[env_id]
    o = om.Oracle(env_id)
    st = om.batch_reset_state(o, 1, seed=0)
    o.set_state(st["qpos"][0], st["qvel"][0], st["ctrl"][0])
    pos = o.field("site_xpos").reshape(-1, 3)
    mat = o.field("site_xmat").reshape(-1, 9)
    for name, ref in g.items():
        sid = o.flat["site_name"].index(name)
        assert np.allclose(pos[sid], ref["pos"], atol=6e-6), (name, pos[sid], ref["pos"])
        if "quat" in ref:
            from gym_kmanip_b200.mjcf import quat_to_mat
            assert np.allclose(mat[sid].reshape(3, 3), quat_to_mat(np.array(ref["quat"])), atol=2e-5)


@pytest.mark.parametrize("env_id", ["KManipSoloArm", "KManipDualArm", "KManipTorso"])
def test_device_ik_algorithm_against_scipy_trf_fixture(env_id):
    """The fixed-iteration projected LM (the algorithm of the CUDA IK, restated in the oracle) lands on the solution
    the reference's scipy TRF finds (ik_mujoco.py:129-135) up to the drift the 2e-6 home regulariser allows along
    the arm's null space, and reaches the same end-effector pose."""
    cases = _load("ik_trf.json")[env_id]
    o = om.Oracle(env_id)
    for c in cases:
        qpos = np.array(c["qpos"])
        o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
        q = o.ik_dls(c["arm"], np.array(c["goal_pos"]), np.array(c["goal_quat"]), qpos)
        assert np.allclose(q, c["q_dls"], atol=1e-12)                 # the oracle reproduces its own fixture
        assert np.abs(q - np.array(c["q_trf"])).max() < 5e-3          # and agrees with scipy's TRF solution
        mask = [o.task.arm_mask[c["arm"]][i] for i in range(o.task.arm_nmask[c["arm"]])]
        r_dls = o.ik_residual(c["arm"], q, c["goal_pos"], c["goal_quat"], qpos[mask])[:6]
        r_trf = o.ik_residual(c["arm"], np.array(c["q_trf"]), c["goal_pos"], c["goal_quat"], qpos[mask])[:6]
        assert np.abs(r_dls - r_trf).max() < 2e-5                     # pose residual (m, and 0.02 * rad)


def test_scipy_trf_live_matches_fixture():
    pytest.importorskip("scipy.optimize")
    cases = _load("ik_trf.json")["KManipSoloArm"][:3]
    o = om.Oracle("KManipSoloArm")
    for c in cases:
        qpos = np.array(c["qpos"])
        o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
        q = o.ik_trf(c["arm"], np.array(c["goal_pos"]), np.array(c["goal_quat"]), qpos)
        assert np.allclose(q, c["q_trf"], atol=1e-7)


@pytest.mark.parametrize("env_id", ["KManipSoloArm", "KManipSoloArmQPos", "KManipDualArm", "KManipDualArmQPos", "KManipTorso"])
def test_oracle_reproduces_golden_trajectory_records(env_id):
    """Teacher-forced: from each stored `before` state + action the oracle must give the stored outputs bit-for-bit
    (same compiler flags, -ffp-contract=off) or to 1e-12 if the libm differs."""
    g = np.load(os.path.join(GOLD, f"traj_{env_id}.npz"))
    o = om.Oracle(env_id)
    n = g["s0_before"].shape[0]
    nq, nv, nu, nm = o.nq, o.nv, o.nu, 7 * o.nmocap
    for t in g["steps"]:
        b = g[f"s{t}_before"]
        st = dict(qpos=b[:, :nq].copy(), qvel=b[:, nq:nq + nv].copy(), ctrl=b[:, nq + nv:nq + nv + nu].copy(),
                  warm=b[:, nq + nv + nu:nq + 2 * nv + nu].copy(), mocap=b[:, nq + 2 * nv + nu:nq + 2 * nv + nu + nm].copy(),
                  time=b[:, nq + 2 * nv + nu + nm].copy(), step=g[f"s{t}_before_step"].copy(), episode=g[f"s{t}_before_episode"].copy())
        if o.nmocap == 0:
            st["mocap"] = np.zeros((n, 0))
        obs, fobs, rew, trunc, flags, ncon, geoms = om.batch_step(o, st, g[f"s{t}_action"], autoreset=True, seed=11)
        assert rel_err(obs, g[f"s{t}_obs"]) < 1e-12
        assert rel_err(rew, g[f"s{t}_reward"]) < 1e-12
        assert rel_err(pack_state(st), g[f"s{t}_after"]) < 1e-12
        assert np.array_equal(trunc, g[f"s{t}_truncated"]) and np.array_equal(ncon, g[f"s{t}_ncon"])
        assert np.array_equal(geoms, g[f"s{t}_geoms"]) and np.array_equal(flags, g[f"s{t}_flags"])
        assert np.array_equal(st["step"], g[f"s{t}_after_step"]) and np.array_equal(st["episode"], g[f"s{t}_after_episode"])
