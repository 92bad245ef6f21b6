"""CPU: the warp-per-env DEVICE code -- km_sim.cuh instantiated with 32 lanes per env, including the register-resident
Newton solver of km_solver_warp.cuh with its shuffles, ballots and votes -- run on the emulated warp of
tests/hostsim/warpemu.h against the oracle.  The emulator steps the 32 lanes as coroutines and checks at every warp
collective that all lanes arrived at the same kind of collective in the same order; a divergent __shfl_sync / __syncwarp
(a hang on the GPU) fails the test with the call sites instead.  Test infrastructure only: no product path runs here."""
import numpy as np
import pytest

import hostsim
from parity_util import comp_rel_err, component_floors, oracle_rollout, rel_err

ENVS = ["KManipSoloArmQPos", "KManipSoloArm", "KManipDualArm", "KManipTorso"]


def _env_state(rec, i):
    return {k: (v[i].copy() if k not in ("time", "step", "episode") else v[i]) for k, v in rec.items()}


@pytest.mark.parametrize("env_id", ENVS)
@pytest.mark.parametrize("dtype,tol_pos,tol_vel", [(64, 1e-10, 1e-10), (32, 2e-5, 2e-3)])
def test_emulated_warp_env_step_matches_oracle(env_id, dtype, tol_pos, tol_vel):
    n = 3
    o, traj = oracle_rollout(env_id, n, 34, seed=5, action_seed=6, round32=(dtype == 32))
    hs = hostsim.HostSim(env_id, dtype, lanes=32)
    worst = dict(pos=0.0, vel=0.0)
    seen_contact = 0
    for rec in traj[::3]:
        for i in range(n):
            st = _env_state(rec["before"], i)
            hs.set_state(st, step=int(st["step"]), episode=int(st["episode"]))
            out = hs.env_step(rec["action"][i], autoreset=True, seed=5, env_id=i)
            assert hs.fault() == "", hs.fault()
            a = hs.get_state()
            if dtype == 32 and rec["ncon_peak"][i] > 0:
                seen_contact += 1
                assert np.abs(a["qpos"] - rec["after"]["qpos"][i]).max() < 1e-3
                continue
            seen_contact += int(rec["ncon"][i] > 0)
            worst["pos"] = max(worst["pos"], rel_err(a["qpos"], rec["after"]["qpos"][i]))
            worst["vel"] = max(worst["vel"], rel_err(a["qvel"], rec["after"]["qvel"][i], floor=1.0))
            if dtype == 64:
                assert out["ncon"] == rec["ncon"][i] and out["flags"] == rec["flags"][i]
    assert worst["pos"] < tol_pos and worst["vel"] < tol_vel, worst
    assert seen_contact > 0          # the table-contact rows of the register-resident solver were exercised


@pytest.mark.parametrize("env_id", ["KManipSoloArmQPos", "KManipDualArm"])
def test_emulated_warp_single_sub_step_fp32_per_component(env_id):
    """km_task.n_sub_steps = 1 (one mj_step per env step), every component on the scale of its own unit: the CPU twin of
    test_gpu_parity.py::test_single_sub_step_parity_fp32_per_component."""
    n = 4
    o, traj = oracle_rollout(env_id, n, 60, seed=7, action_seed=8, round32=True, n_sub_steps=1)
    hs = hostsim.HostSim(env_id, 32, lanes=32, n_sub_steps=1)
    fq, fv = component_floors(o.flat, o.nq, o.nv)
    wp = wv = 0.0
    for rec in traj[::2]:
        for i in range(n):
            if rec["ncon_peak"][i] > 0:
                continue
            st = _env_state(rec["before"], i)
            hs.set_state(st, step=int(st["step"]), episode=int(st["episode"]))
            hs.env_step(rec["action"][i], autoreset=True, seed=7, env_id=i)
            assert hs.fault() == "", hs.fault()
            a = hs.get_state()
            wp = max(wp, comp_rel_err(a["qpos"][None], rec["after"]["qpos"][i][None], fq))
            wv = max(wv, comp_rel_err(a["qvel"][None], rec["after"]["qvel"][i][None], fv))
    print(env_id, "one fp32 sub-step, free envs: pos %.2e vel %.2e" % (wp, wv))
    assert wp < 1e-5 and wv < 5e-4      # see test_gpu_parity.py::test_single_sub_step_parity_fp32_per_component


@pytest.mark.parametrize("env_id", ["KManipSoloArm", "KManipDualArm", "KManipTorso"])
def test_emulated_warp_finger_pad_contacts_match_oracle(env_id):
    """Coupled case on the warp-per-env code: the cube is placed against the finger pads (in half of the cases on the table
    too), so pad contacts couple the arm and cube blocks and the register-resident solver takes its dense path (up to
    four contacts; beyond that the generic solver)."""
    from oracle import oracle as om
    o = om.Oracle(env_id)
    hs = hostsim.HostSim(env_id, 64, lanes=32)
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
        assert hs.fault() == "", hs.fault()
        o.set_state(state["qpos"], state["qvel"], state["ctrl"], state["warm"], 0.0, state["mocap"] if o.nmocap else None)
        obs, rew = o.step(act)
        a, b = hs.get_state(), o.get_state()
        assert rel_err(a["qpos"], b["qpos"]) < 1e-10 and rel_err(a["qvel"], b["qvel"], floor=1.0) < 1e-9, (trial, rel_err(a["qvel"], b["qvel"]))
        assert rel_err(out["obs"], obs, floor=1.0) < 1e-9
    assert seen >= 4


def test_emulator_reports_divergent_collectives():
    """The checker itself: a collective reached by only some lanes is reported, not silently executed."""
    import ctypes as C
    L = hostsim.lib()
    L.hs_emu_selftest.restype = C.c_int
    assert L.hs_emu_selftest(0) == 0       # uniform collectives: fine
    assert L.hs_emu_selftest(1) == 1       # half of the lanes skip a shuffle: reported
