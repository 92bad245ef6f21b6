"""CPU: physical self-consistency of the oracle (the substitute for golden vectors the reference does not have,
SURVEY.md section 4): mass matrix symmetric positive definite, gravity torques = dU/dq, equations of motion closed by
the constraint forces, complementarity of the constraint rows, analytic Jacobians vs finite differences."""
import numpy as np
import pytest

from oracle import oracle as om

SCENES = ["KManipSoloArm", "KManipDualArm", "KManipTorso"]


def _random_state(o, rng, vel=1.0, spread=0.3):
    st = om.batch_reset_state(o, 1, seed=int(rng.integers(1 << 30)))
    qpos = st["qpos"][0].copy()
    rngs = np.array(o.flat["jnt_range"])
    for j in range(o.nu):
        lo, hi = rngs[j]
        qpos[j] = np.clip(qpos[j] + rng.uniform(-spread, spread) * (hi - lo), lo + 1e-3 * (hi - lo), hi - 1e-3 * (hi - lo))
    q = rng.normal(size=4)
    qpos[-4:] = q / np.linalg.norm(q)
    qvel = rng.normal(size=o.nv) * vel
    return qpos, qvel


@pytest.mark.parametrize("env_id", SCENES)
def test_mass_matrix_symmetric_positive_definite(env_id):
    o = om.Oracle(env_id)
    rng = np.random.default_rng(0)
    for _ in range(5):
        qpos, qvel = _random_state(o, rng)
        o.set_state(qpos, qvel, qpos[: o.nu])
        M = o.field("qM").reshape(o.nv, o.nv)
        assert np.allclose(M, M.T, atol=1e-14)
        assert np.linalg.eigvalsh(M).min() > 1e-6


@pytest.mark.parametrize("env_id", SCENES)
def test_gravity_bias_is_gradient_of_potential_energy(env_id):
    """At zero velocity the RNE bias force is dU/dq with U = sum_b m_b g z_com_b (independent of the RNE code path:
    it only uses the forward kinematics of the body COMs)."""
    o = om.Oracle(env_id)
    rng = np.random.default_rng(1)
    mass = np.array(o.flat["body_mass"])
    g = -np.array(o.flat["opt"]["gravity"])[2]

    def U(qpos):
        o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
        return float(np.sum(mass * g * o.field("xipos").reshape(-1, 3)[:, 2]))

    qpos, _ = _random_state(o, rng)
    o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
    bias = o.field("qfrc_bias")
    eps = 1e-6
    for j in range(o.nu):           # articulated dofs: qpos address == dof address
        qp, qm = qpos.copy(), qpos.copy()
        qp[j] += eps
        qm[j] -= eps
        fd = (U(qp) - U(qm)) / (2 * eps)
        assert abs(fd - bias[j]) < 1e-6 * max(1.0, abs(bias[j])), (j, fd, bias[j])
    # free cube: bias = [0, 0, m g, 0, 0, 0]
    cube_m = mass[o.task.cube_body]
    assert np.allclose(bias[o.nu:], [0, 0, cube_m * g, 0, 0, 0], atol=1e-12)


@pytest.mark.parametrize("env_id", SCENES)
def test_equations_of_motion_and_constraint_rows(env_id):
    """After the forward pass: M qacc = qfrc_smooth + J^T f, qfrc_constraint = J^T f, and every row's force obeys its
    own law at the solution (friction loss: |f| <= loss; limits and pyramid edges: f >= 0, f = -D min(jar, 0))."""
    o = om.Oracle(env_id)
    rng = np.random.default_rng(2)
    seen_contact = False
    for trial in range(12):
        qpos, qvel = _random_state(o, rng, vel=0.5)
        if trial % 2 == 0:      # put the cube on the table (tilted: fewer than four corners touch) to exercise contact rows
            qpos[-7:-4] = [0.2, 0.6, 0.5 + 0.0199]
            qpos[-4:] = [1, 0, 0, 0] if trial % 4 == 0 else [0.9998, 0.02, 0.0, 0.0]
            qpos[-4:] /= np.linalg.norm(qpos[-4:])
            qvel[-6:] *= 0.01
        o.set_state(qpos, qvel, qpos[: o.nu])
        o.mj_forward()
        nv = o.nv
        M = o.field("qM").reshape(nv, nv)
        qacc, smooth, qfc = o.field("qacc"), o.field("qfrc_smooth"), o.field("qfrc_constraint")
        nefc = int(o.field("nefc")[0])
        J = o.field("efc_J").reshape(nefc, nv)
        f, aref, Dr = o.field("efc_force"), o.field("efc_aref"), o.field("efc_D")
        typ, loss = o.field("efc_type").astype(int), o.field("efc_frictionloss")
        assert np.allclose(qfc, J.T @ f, atol=1e-9 * max(1.0, np.abs(f).max()))
        scale = max(1.0, np.abs(smooth).max(), np.abs(qfc).max())
        assert np.abs(M @ qacc - smooth - qfc).max() < 1e-6 * scale      # Newton stops at tolerance 1e-8 (scaled)
        jar = J @ qacc - aref
        for r in range(nefc):
            if loss[r] > 0:                                            # friction-loss row
                assert abs(f[r]) <= loss[r] * (1 + 1e-12)
                if abs(f[r]) < loss[r] * (1 - 1e-9):
                    assert abs(f[r] + Dr[r] * jar[r]) < 1e-7 * max(1.0, abs(f[r]))
            else:                                                      # limit or pyramidal contact edge
                assert f[r] >= 0
                assert abs(f[r] + Dr[r] * min(jar[r], 0.0)) < 1e-7 * max(1.0, abs(f[r]))
        seen_contact = seen_contact or int(o.field("ncon")[0]) > 0
    assert seen_contact


@pytest.mark.parametrize("env_id", SCENES)
def test_ik_jacobian_pose_rows_match_finite_differences(env_id):
    """ik_jac's pose rows (ik_mujoco.py:56-97) are the derivative of ik_res's pose rows (ik_mujoco.py:20-53); the
    regulariser rows are deliberately NOT each other's derivative in the reference."""
    o = om.Oracle(env_id)
    rng = np.random.default_rng(4)
    from scipy.spatial.transform import Rotation as R
    for arm in range(o.task.n_arm):
        qpos, _ = _random_state(o, rng, spread=0.1)
        o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
        n = o.task.arm_nmask[arm]
        mask = [o.task.arm_mask[arm][i] for i in range(n)]
        sid = o.task.arm_site[arm]
        gp = o.field("site_xpos").reshape(-1, 3)[sid] + rng.uniform(-0.01, 0.01, 3)
        rot = R.from_matrix(o.field("site_xmat").reshape(-1, 3, 3)[sid]) * R.from_rotvec(rng.uniform(-0.1, 0.1, 3))
        gq = rot.as_quat()[[3, 0, 1, 2]]
        x = qpos[mask]
        Jm = o.ik_jacobian(arm, x, gq)
        eps = 1e-6
        for c in range(n):
            xp, xm = x.copy(), x.copy()
            xp[c] += eps
            xm[c] -= eps
            fd = (o.ik_residual(arm, xp, gp, gq, x) - o.ik_residual(arm, xm, gp, gq, x)) / (2 * eps)
            assert np.allclose(Jm[:6, c], fd[:6], atol=2e-8), (arm, c, Jm[:6, c], fd[:6])
        # the regulariser blocks are constant diagonals with the reference's (mismatched) weights
        assert np.allclose(Jm[6:6 + n], 9e-3 * np.eye(n)) and np.allclose(Jm[6 + n:], 9e-3 * np.eye(n))


def test_free_fall_of_the_cube_is_semi_implicit_euler():
    """Cube above the table, no contact: z follows the semi-implicit Euler recursion with the friction-loss row of
    its free joint saturated against the motion (0.01 N, scene.xml:19-20 + completion spec)."""
    o = om.Oracle("KManipSoloArmQPos")
    st = om.batch_reset_state(o, 1, seed=0)
    qpos = st["qpos"][0].copy()
    qpos[-7:-4] = [0.2, 0.6, 0.9]
    o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
    h = o.flat["opt"]["timestep"]
    g = -o.flat["opt"]["gravity"][2]
    m = o.flat["body_mass"][o.task.cube_body]
    loss = o.flat["dof_frictionloss"][o.nu + 2]
    z, v = 0.9, 0.0
    for _ in range(20):
        o.mj_step()
        v += h * (-g + loss / m)
        z += h * v
    s = o.get_state()
    assert abs(s["qpos"][-5] - z) < 1e-9 and abs(s["qvel"][o.nu + 2] - v) < 1e-8


def test_reward_terms_follow_env_sim():
    """get_reward (env_sim.py:148-161): -0.01 |qvel| + 0.01 / (|cube - ee body| + 1e-6) per gripper in act_list."""
    o = om.Oracle("KManipDualArm")
    rng = np.random.default_rng(9)
    qpos, qvel = _random_state(o, rng)
    o.set_state(qpos, qvel, qpos[: o.nu])
    r = o.reward()
    xpos = o.field("xpos").reshape(-1, 3)
    expect = -0.01 * np.linalg.norm(qvel)
    for a in range(o.task.n_arm):
        expect += 0.01 / (np.linalg.norm(xpos[o.task.cube_body] - xpos[o.task.arm_eebody[a]]) + 1e-6)
    assert abs(r - expect) < 1e-12


def test_observation_normalisation_follows_env_sim():
    """get_observation (env_sim.py:110-139): q_pos scaled by jnt_range, q_vel by pi, cube_pos by the spawn range."""
    o = om.Oracle("KManipSoloArm")
    rng = np.random.default_rng(10)
    qpos, qvel = _random_state(o, rng, vel=2.0)
    o.set_state(qpos, qvel, qpos[: o.nu])
    obs = o.obs()
    ql = o.task.q_len
    rngs = np.array(o.flat["jnt_range"])[:ql]
    assert np.allclose(obs[:ql], np.clip((qpos[:ql] - rngs[:, 0]) / (rngs[:, 1] - rngs[:, 0]), -1, 1))
    assert np.allclose(obs[ql:2 * ql], np.clip(qvel[:ql] / np.pi, -1, 1))
    lo, hi = np.array([0.1, 0.5, 0.6]), np.array([0.3, 0.7, 0.7])
    assert np.allclose(obs[2 * ql:2 * ql + 3], np.clip((qpos[-7:-4] - lo) / (hi - lo), -1, 1))
    assert np.allclose(obs[2 * ql + 3:], qpos[-4:])
