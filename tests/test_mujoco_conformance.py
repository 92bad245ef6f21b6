"""Conformance with the real engine, wherever it is installed.

``mujoco`` cannot be installed in the authoring image or on the GPU boxes (no network, no wheel), so the second half
of this file SKIPS there with that reason; it runs unchanged on any machine with ``pip install mujoco``.

What runs everywhere (CPU): the completed model of every scene is exported as a self-contained MJCF
(gym_kmanip_b200/mjcf_export.py: primitives, explicit <inertial>s and <contact><pair>s -- no meshes), the document is
well-formed and carries the same bodies / joints / geoms / pairs / actuators in the same order as the flat model.

What runs with mujoco: (1) ``mujoco.MjModel.from_xml_string`` compiles that MJCF; the flat model filled from the
MjModel (mjcf_export.flat_from_mjmodel -- the path INTEGRATION.md 1 describes) agrees with the one mjcf.py builds,
including the compile-time constants dof_invweight0 / body_invweight0 / stat.meaninertia that mjcf.py restates; (2) the
oracle, built on the MjModel-derived arrays, is compared with ``mj_step1`` / ``mj_step2`` sub-step by sub-step from
identical states (reference path: gym_kmanip/env_sim.py:206-211 ``mujoco.Physics.from_xml_path`` + ``physics.step``):
contact counts, qacc, qpos, qvel.  That is the pin the oracle otherwise lacks ("parity unpinned").
"""
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from gym_kmanip_b200 import constants as K, mjcf, mjcf_export

SCENES = ["solo_arm", "dual_arm", "torso"]
ENV_OF = {"solo_arm": "KManipSoloArm", "dual_arm": "KManipDualArm", "torso": "KManipTorso"}


@pytest.mark.parametrize("scene", SCENES)
def test_completed_model_exports_as_wellformed_mjcf(scene):
    flat = mjcf.load_flat(scene)
    root = ET.fromstring(mjcf_export.completed_mjcf(flat))
    bodies = [b.attrib["name"] for b in root.find("worldbody").iter("body")]
    assert bodies == flat["body_name"][1:]                              # depth-first order = body ids
    joints = [j.attrib["name"] for j in root.find("worldbody").iter("joint")]
    assert joints == flat["jnt_name"]
    assert [g.attrib["name"] for g in root.find("worldbody").iter("geom")] == flat["geom_name"]
    assert len(root.find("contact").findall("pair")) == flat["npair"] and len(root.find("actuator")) == flat["nu"]
    assert sum(1 for b in root.find("worldbody").iter("body") if b.attrib.get("mocap") == "true") == flat["nmocap"]
    inert = {b.attrib["name"]: b.find("inertial") for b in root.find("worldbody").iter("body")}
    for b in range(1, flat["nbody"]):
        assert (inert[flat["body_name"][b]] is not None) == (flat["body_mass"][b] > 0)
    assert float(root.find("option").attrib["timestep"]) == flat["opt"]["timestep"]


# ------------------------------------------------------------------------------------------------ needs the real engine
def _mj():
    return pytest.importorskip("mujoco", reason="mujoco is not installable in this image (no network / wheel); "
                                                "run this file where `pip install mujoco` works")


@pytest.mark.parametrize("scene", SCENES)
def test_flat_model_matches_mjmodel(scene):
    mujoco = _mj()
    flat = mjcf.load_flat(scene)
    m = mujoco.MjModel.from_xml_string(mjcf_export.completed_mjcf(flat))
    got = mjcf_export.flat_from_mjmodel(m, template=flat)
    for k in ("nbody", "njnt", "nq", "nv", "nu", "nsite", "ngeom", "npair", "nmocap"):
        assert got[k] == flat[k], k
    assert got["body_name"] == flat["body_name"] and got["jnt_name"] == flat["jnt_name"] and got["geom_name"] == flat["geom_name"]
    for k in mjcf_export.NUMERIC_KEYS:
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(flat[k], dtype=np.float64)
        if k == "geom_size":      # MuJoCo keeps all three size slots; only the ones the primitive uses are defined
            a, b = a.copy(), b.copy()
            for g, t in enumerate(flat["geom_type"]):
                if t == 2:
                    a[g, 1:] = b[g, 1:] = 0
        assert a.shape == b.shape, k
        assert np.allclose(a, b, rtol=1e-9, atol=1e-12), (k, np.abs(a - b).max())
    # the compile-time constants mjcf.py restates: (M^-1)_ii, tr(J M^-1 J^T)/3, mean(diag M)
    assert np.allclose(got["dof_invweight0"], flat["dof_invweight0"], rtol=1e-8)
    assert np.allclose(got["body_invweight0"], flat["body_invweight0"], rtol=1e-8, atol=1e-12)
    assert abs(got["meaninertia"] - flat["meaninertia"]) < 1e-9 * flat["meaninertia"]


@pytest.mark.parametrize("scene", SCENES)
def test_oracle_matches_mj_step_per_sub_step(scene):
    """Teacher-forced: the oracle's states along a random-action rollout (cube falling, landing, resting) are handed to
    MuJoCo; one mj_step1 + mj_step2 on both sides from each state."""
    mujoco = _mj()
    from oracle import oracle as om
    from parity_util import oracle_rollout, rel_err
    flat = mjcf.load_flat(scene)
    m = mujoco.MjModel.from_xml_string(mjcf_export.completed_mjcf(flat))
    d = mujoco.MjData(m)
    env_id = ENV_OF[scene]
    o = om.Oracle(env_id, flat=mjcf_export.flat_from_mjmodel(m, template=flat))     # the oracle on MuJoCo's own arrays
    _, traj = oracle_rollout(env_id, 8, 40, seed=4, action_seed=9)
    worst = dict(qacc=0.0, qpos=0.0, qvel=0.0)
    for rec in traj[::3]:
        b = rec["before"]
        for i in range(8):
            qpos, qvel, ctrl, warm = b["qpos"][i], b["qvel"][i], b["ctrl"][i], b["warm"][i]
            mujoco.mj_resetData(m, d)
            d.qpos[:], d.qvel[:], d.ctrl[:], d.qacc_warmstart[:], d.time = qpos, qvel, ctrl, warm, float(b["time"][i])
            if m.nmocap:
                d.mocap_pos[:] = b["mocap"][i].reshape(m.nmocap, 7)[:, :3]
                d.mocap_quat[:] = b["mocap"][i].reshape(m.nmocap, 7)[:, 3:]
            o.set_state(qpos, qvel, ctrl, warm, float(b["time"][i]), b["mocap"][i] if o.nmocap else None)
            mujoco.mj_step1(m, d)
            o.mj_step1()
            assert d.ncon == int(o.field("ncon")[0])
            assert d.nefc == int(o.field("nefc")[0])
            mujoco.mj_step2(m, d)
            o.mj_step2()
            s = o.get_state()
            worst["qacc"] = max(worst["qacc"], rel_err(o.field("qacc"), d.qacc, floor=1.0))
            worst["qpos"] = max(worst["qpos"], rel_err(s["qpos"], d.qpos))
            worst["qvel"] = max(worst["qvel"], rel_err(s["qvel"], d.qvel, floor=1.0))
    # the Newton solver stops at tolerance 1e-8 on both sides; their iterates need not coincide beyond that
    assert worst["qacc"] < 1e-6 and worst["qpos"] < 1e-9 and worst["qvel"] < 1e-8, worst


def test_env_step_matches_dm_control_order():
    """dm_control's legacy step order (mj_step2; mj_step x 9; mj_step1) through mujoco alone: ten sub-steps of the oracle's
    env step against the same sequence of mj_* calls, joint-position actions (no IK side effects)."""
    mujoco = _mj()
    from oracle import oracle as om
    from parity_util import rel_err
    flat = mjcf.load_flat("solo_arm")
    m = mujoco.MjModel.from_xml_string(mjcf_export.completed_mjcf(flat))
    d = mujoco.MjData(m)
    o = om.Oracle("KManipSoloArmQPos", flat=mjcf_export.flat_from_mjmodel(m, template=flat))
    xyz = np.array([0.2, 0.6, 0.65])
    o.reset(xyz)
    s = o.get_state()
    mujoco.mj_resetData(m, d)
    d.qpos[:], d.qvel[:], d.ctrl[:] = s["qpos"], s["qvel"], s["ctrl"]
    mujoco.mj_forward(m, d)
    rng = np.random.default_rng(0)
    for t in range(12):
        a = rng.uniform(-1, 1, o.task.act_dim).astype(np.float32)
        o.before_step(a)
        d.ctrl[:] = o.get_state()["ctrl"]
        o.step(a)
        mujoco.mj_step2(m, d)
        for _ in range(int(round(K.CONTROL_TIMESTEP / m.opt.timestep)) - 1):
            mujoco.mj_step(m, d)
        mujoco.mj_step1(m, d)
        s = o.get_state()
        assert rel_err(s["qpos"], d.qpos) < 1e-7 and rel_err(s["qvel"], d.qvel, floor=1.0) < 1e-6, t
