"""CPU: the kernel source (gym_kmanip_b200/csrc/km_sim.cuh) built for the host with one lane per env
(tests/hostsim, test infrastructure only) against the oracle, stage by stage and per env step, fp64 and fp32.
This is how the simulator source is checked on machines without a GPU; the GPU build of the same source is
checked through the C-ABI in test_gpu_parity.py."""
import numpy as np
import pytest

import hostsim
from oracle import oracle as om
from parity_util import CONTACT_TOL_POS_F32, CONTACT_TOL_VEL_F32, oracle_rollout, rel_err

ENVS = ["KManipSoloArmQPos", "KManipSoloArm", "KManipDualArm", "KManipTorso"]


def _env_state(rec, i):
    return {k: (v[i].copy() if k not in ("time", "step", "episode") else v[i]) for k, v in rec.items()}


@pytest.mark.parametrize("env_id", ENVS)
def test_sub_step_stages_match_oracle_fp64(env_id):
    """mj_step1 products (FK, mass matrix, bias, constraint reference) and mj_step2 products (smooth and constrained
    accelerations, constraint forces, integrated state) from identical states."""
    o, traj = oracle_rollout(env_id, 4, 12, seed=2, action_seed=4)
    hs = hostsim.HostSim(env_id, 64)
    for rec in traj[3::4]:
        for i in range(4):
            st = _env_state(rec["before"], i)
            o.set_state(st["qpos"], st["qvel"], st["ctrl"], st["warm"], st["time"], st["mocap"])
            hs.set_state(st)
            hs.step1()
            assert rel_err(hs.field("qM"), o.field("qM")) < 1e-12
            assert rel_err(hs.field("qfrc_bias"), o.field("qfrc_bias")) < 1e-11
            assert int(hs.field("ncon")[0]) == int(o.field("ncon")[0])
            assert int(hs.field("nefc")[0]) == int(o.field("nefc")[0])
            assert rel_err(hs.field("efc_aref"), o.field("efc_aref")) < 1e-10
            assert rel_err(hs.field("efc_D"), o.field("efc_D")) < 1e-12
            assert rel_err(hs.field("efc_J"), o.field("efc_J")) < 1e-12
            if int(o.field("ncon")[0]):
                assert rel_err(hs.field("contact_dist"), o.field("contact_dist"), floor=1e-9) < 1e-8
                assert rel_err(hs.field("contact_frame"), o.field("contact_frame")) < 1e-12
            o.mj_step2()
            hs.step2()
            assert rel_err(hs.field("qacc_smooth"), o.field("qacc_smooth")) < 1e-10
            assert rel_err(hs.field("qacc"), o.field("qacc")) < 1e-9
            # constraint force in joint space, from the equation of motion (the thread-per-env solver keeps no row forces)
            qfc = hs.field("qM").reshape(hs.nv, hs.nv) @ hs.field("qacc") - hs.field("qfrc_smooth")
            assert rel_err(qfc, o.field("qfrc_constraint"), floor=1.0) < 1e-8
            a, b = hs.get_state(), o.get_state()
            assert rel_err(a["qpos"], b["qpos"]) < 1e-12 and rel_err(a["qvel"], b["qvel"]) < 1e-11


@pytest.mark.parametrize("env_id", ENVS)
@pytest.mark.parametrize("dtype,tol_pos,tol_vel", [(64, 1e-10, 1e-10), (32, 2e-5, 2e-3)])
def test_env_step_matches_oracle(env_id, dtype, tol_pos, tol_vel):
    """Whole env step (action decode + IK + 10 sub-steps + reward / observation / truncation) from identical states.
    fp32: the teacher states are float32-representable (what the fp32 build can be handed at all); envs whose cube
    touches something are held to the looser contact tolerance of parity_util (DESIGN.md "fp32 and the cube")."""
    o, traj = oracle_rollout(env_id, 6, 66, seed=5, action_seed=6, round32=(dtype == 32))
    hs = hostsim.HostSim(env_id, dtype)
    worst = dict(pos=0.0, vel=0.0, obs=0.0, cpos=0.0, cvel=0.0)
    for rec in traj[::5] + traj[62:66]:
        for i in range(6):
            st = _env_state(rec["before"], i)
            hs.set_state(st, step=int(st["step"]), episode=int(st["episode"]))
            out = hs.env_step(rec["action"][i], autoreset=True, seed=5, env_id=i)
            a = hs.get_state()
            ep = max(rel_err(a["qpos"], rec["after"]["qpos"][i]), rel_err(a["ctrl"], rec["after"]["ctrl"][i]))
            ev = rel_err(a["qvel"], rec["after"]["qvel"][i], floor=1.0)
            eo = max(rel_err(out["obs"], rec["obs"][i], floor=1.0), rel_err(out["reward"], rec["reward"][i], floor=1.0))
            touching = dtype == 32 and rec["ncon_peak"][i] > 0
            if touching:
                worst["cpos"], worst["cvel"] = max(worst["cpos"], ep, eo), max(worst["cvel"], ev)
            else:
                worst["pos"], worst["vel"], worst["obs"] = max(worst["pos"], ep), max(worst["vel"], ev), max(worst["obs"], eo)
            assert out["truncated"] == bool(rec["truncated"][i])
            assert a["step"] == rec["after"]["step"][i] and a["episode"] == rec["after"]["episode"][i]
            if dtype == 64:
                assert out["ncon"] == rec["ncon"][i]
                assert np.array_equal(out["geoms"], rec["geoms"][i][: len(out["geoms"])])
                assert out["flags"] == rec["flags"][i]
    assert worst["pos"] < tol_pos and worst["vel"] < tol_vel and worst["obs"] < max(tol_vel, 10 * tol_pos), worst
    assert worst["cpos"] < CONTACT_TOL_POS_F32 and worst["cvel"] < CONTACT_TOL_VEL_F32, worst


def test_reset_matches_oracle_spawn():
    """initialize_episode (env_sim.py:23-36) with the counter-based spawn: same cube positions for the same
    (seed, global env id, episode) on both sides."""
    o = om.Oracle("KManipSoloArm")
    hs = hostsim.HostSim("KManipSoloArm", 64)
    for env_id in (0, 1, 77, 2 ** 33 + 5):
        for ep in (0, 3):
            st = dict(qpos=np.zeros(hs.nq), qvel=np.zeros(hs.nv), ctrl=np.zeros(hs.nu), warm=np.zeros(hs.nv),
                      mocap=np.zeros(7 * hs.nmocap), time=0.0)
            hs.set_state(st, step=9, episode=ep)
            obs = hs.reset(seed=42, env_id=env_id)
            xyz = o.spawn(42, env_id, ep)
            assert np.allclose(hs.get_state()["qpos"][-7:-4], xyz, atol=1e-15)
            assert np.allclose(obs, o.reset(xyz), atol=1e-14)
            assert 0.1 <= xyz[0] <= 0.3 and 0.5 <= xyz[1] <= 0.7 and 0.6 <= xyz[2] <= 0.7


@pytest.mark.parametrize("env_id", ["KManipSoloArm", "KManipDualArm", "KManipTorso"])
def test_finger_pad_contacts_match_oracle(env_id):
    """Coupled case: the cube is placed against the finger pads (and, in half of the cases, on the table too) so that
    pad contacts couple the arm and cube blocks of the solver's Hessian."""
    o = om.Oracle(env_id)
    hs = hostsim.HostSim(env_id, 64)
    st0 = om.batch_reset_state(o, 1, seed=1)
    rng = np.random.default_rng(3)
    pads = [i for i, nm in enumerate(o.flat["geom_name"]) if nm.startswith("finger_pad")]
    seen = 0
    for trial in range(8):
        qpos = st0["qpos"][0].copy()
        qpos[: o.nu] += rng.uniform(-0.05, 0.05, o.nu) * (np.arange(o.nu) < o.nu)
        rngs = np.array(o.flat["jnt_range"])[: o.nu]
        qpos[: o.nu] = np.clip(qpos[: o.nu], rngs[:, 0] + 1e-3, rngs[:, 1] - 1e-3)
        o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
        gx = o.field("geom_xpos").reshape(-1, 3)
        p = gx[pads[trial % len(pads)]]
        # cube face 2 mm inside the pad sphere (radius 0.01), approached along a random axis
        ax = trial % 3
        off = np.zeros(3)
        off[ax] = (0.02 + 0.01 - 0.002) * (1 if trial % 2 else -1)
        qpos[-7:-4] = p + off
        qpos[-4:] = [1, 0, 0, 0]
        qvel = rng.normal(size=o.nv) * 0.1
        state = dict(qpos=qpos, qvel=qvel, ctrl=qpos[: o.nu].copy(), warm=np.zeros(o.nv), mocap=st0["mocap"][0][: 7 * o.nmocap].copy(), time=0.0)
        o.set_state(state["qpos"], state["qvel"], state["ctrl"], state["warm"], 0.0, state["mocap"] if o.nmocap else None)
        r, fl = o.reward(with_flags=True)
        if not fl & 6:
            continue
        seen += 1
        act = rng.uniform(-1, 1, o.task.act_dim).astype(np.float32)
        hs.set_state(state)
        out = hs.env_step(act)
        o.set_state(state["qpos"], state["qvel"], state["ctrl"], state["warm"], 0.0, state["mocap"] if o.nmocap else None)
        obs, rew = o.step(act)
        a, b = hs.get_state(), o.get_state()
        assert rel_err(a["qpos"], b["qpos"]) < 1e-10 and rel_err(a["qvel"], b["qvel"], floor=1.0) < 1e-9, (trial, rel_err(a["qvel"], b["qvel"]))
        assert rel_err(out["obs"], obs, floor=1.0) < 1e-9
    assert seen >= 4
