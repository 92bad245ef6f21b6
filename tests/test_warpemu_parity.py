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


def test_emulator_reports_divergent_collectives():
    """The checker itself: a collective reached by only some lanes is reported, not silently executed."""
    import ctypes as C
    L = hostsim.lib()
    L.hs_emu_selftest.restype = C.c_int
    assert L.hs_emu_selftest(0) == 0       # uniform collectives: fine
    assert L.hs_emu_selftest(1) == 1       # half of the lanes skip a shuffle: reported
