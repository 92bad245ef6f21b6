"""ctypes front-end of the CPU oracle (oracle/kmanip_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never imported by gym_kmanip_b200.

PARITY UNPINNED (see oracle/ko_model.h): MuJoCo / dm_control are not importable here; scipy is, and
``ik_mode="trf"`` drives the restated ik_res / ik_jac with the real scipy.optimize.least_squares exactly
as reference gym_kmanip/ik_mujoco.py:129-135 does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from gym_kmanip_b200 import constants as K          # noqa: E402
from gym_kmanip_b200 import flatmodel, mjcf         # noqa: E402

_LIBS: Dict[str, C.CDLL] = {}
PY_IK = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))


def build(force: bool = False) -> None:
    so = os.path.join(_HERE, "_build", "libkmanip_oracle.so")
    src = os.path.join(_HERE, "kmanip_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib(flops: bool = False) -> C.CDLL:
    key = "flops" if flops else "plain"
    if key not in _LIBS:
        build()
        name = "libkmanip_oracle_flops.so" if flops else "libkmanip_oracle.so"
        L = C.CDLL(os.path.join(_HERE, "_build", name))
        L.ko_data_new.restype = C.c_void_p
        L.ko_data_new.argtypes = [C.c_void_p]
        L.ko_data_free.argtypes = [C.c_void_p]
        L.ko_get_reward.restype = C.c_double
        L.ko_flops_get.restype = C.c_ulonglong
        L.ko_get_field.restype = C.c_int
        _LIBS[key] = L
    return _LIBS[key]


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """One simulated env on the CPU, mirroring KManipTask + dm_control's episode loop."""

    def __init__(self, env_id: str = "KManipSoloArm", flat: Optional[dict] = None, ik_mode: str = "dls",
                 ik_iters: int = K.DEVICE_IK_ITERS, ik_teleport: bool = True, flops: bool = False, n_sub_steps: int = 0, **opt):
        self.kw = K.ENV_REGISTRY[env_id]
        self.flat = dict(flat if flat is not None else mjcf.load_flat(mjcf.scene_of_mjcf(self.kw["mjcf_filename"])))
        if opt:
            self.flat["opt"] = dict(self.flat["opt"], **opt)
        self.pm = flatmodel.PackedModel(self.flat)
        self.task = flatmodel.make_task(self.flat, self.kw, ik_iters=ik_iters, ik_teleport=ik_teleport, n_sub_steps=n_sub_steps)
        self.L = lib(flops)
        self.d = C.c_void_p(self.L.ko_data_new(self.pm.ref()))
        self.nq, self.nv, self.nu = self.flat["nq"], self.flat["nv"], self.flat["nu"]
        self.nmocap = self.flat["nmocap"]
        self.obs_dim = self.L.ko_obs_size(C.byref(self.task))
        self.ik_mode = ik_mode
        self.ik_nfev = []
        self._cb = PY_IK(self._ik_trf) if ik_mode == "trf" else None

    def __del__(self):
        try:
            self.L.ko_data_free(self.d)
        except Exception:
            pass

    # ---- state
    def set_state(self, qpos, qvel, ctrl, warm=None, time=0.0, mocap=None):
        qpos, qvel, ctrl = (np.ascontiguousarray(x, dtype=np.float64) for x in (qpos, qvel, ctrl))
        warm = np.zeros(self.nv) if warm is None else np.ascontiguousarray(warm, dtype=np.float64)
        mp = None if mocap is None else _dp(np.ascontiguousarray(mocap, dtype=np.float64))
        self.L.ko_set_state(self.pm.ref(), self.d, _dp(qpos), _dp(qvel), _dp(ctrl), _dp(warm), C.c_double(time), mp)

    def get_state(self):
        qpos, qvel, ctrl, warm = np.zeros(self.nq), np.zeros(self.nv), np.zeros(self.nu), np.zeros(self.nv)
        mocap = np.zeros(7 * max(self.nmocap, 1))
        t = C.c_double(0)
        self.L.ko_get_state(self.pm.ref(), self.d, _dp(qpos), _dp(qvel), _dp(ctrl), _dp(warm), C.byref(t), _dp(mocap))
        return dict(qpos=qpos, qvel=qvel, ctrl=ctrl, warm=warm, time=t.value, mocap=mocap[: 7 * self.nmocap])

    def field(self, name: str, cap: int = 65536) -> np.ndarray:
        out = np.zeros(cap)
        n = self.L.ko_get_field(self.pm.ref(), self.d, name.encode(), _dp(out), cap)
        if n < 0:
            raise KeyError(name)
        return out[:n].copy()

    # ---- episode API
    def spawn(self, seed: int, env_id: int, episode: int) -> np.ndarray:
        xyz = np.zeros(3)
        self.L.ko_spawn(C.byref(self.task), C.c_ulonglong(seed), C.c_ulonglong(env_id), C.c_uint(episode), _dp(xyz))
        return xyz

    def reset(self, cube_xyz):
        xyz = np.ascontiguousarray(cube_xyz, dtype=np.float64)
        self.L.ko_reset(self.pm.ref(), self.d, C.byref(self.task), _dp(xyz))
        return self.obs()

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float32)
        assert a.size == self.task.act_dim
        self.L.ko_env_step(self.pm.ref(), self.d, C.byref(self.task), a.ctypes.data_as(C.POINTER(C.c_float)), self._cb)
        return self.obs(), self.reward()

    def before_step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float32)
        self.L.ko_before_step_only(self.pm.ref(), self.d, C.byref(self.task), a.ctypes.data_as(C.POINTER(C.c_float)), self._cb)

    def mj_step(self):
        self.L.ko_mj_step(self.pm.ref(), self.d)

    def mj_step1(self):
        self.L.ko_mj_step1(self.pm.ref(), self.d)

    def mj_step2(self):
        self.L.ko_mj_step2(self.pm.ref(), self.d)

    def mj_forward(self, disable_actuation: bool = False):
        self.L.ko_mj_forward(self.pm.ref(), self.d, int(disable_actuation))

    def obs(self) -> np.ndarray:
        o = np.zeros(self.obs_dim)
        self.L.ko_get_obs(self.pm.ref(), self.d, C.byref(self.task), _dp(o))
        return o

    def reward(self, with_flags: bool = False):
        fl = C.c_int(0)
        r = self.L.ko_get_reward(self.pm.ref(), self.d, C.byref(self.task), C.byref(fl))
        return (r, fl.value) if with_flags else r

    # ---- IK pieces
    def ik_residual(self, arm, q, goal_pos, goal_quat, q_prev_mask):
        n = self.task.arm_nmask[arm]
        res = np.zeros(6 + 2 * n)
        self.L.ko_ik_residual(self.pm.ref(), self.d, C.byref(self.task), arm, _dp(np.ascontiguousarray(q, dtype=np.float64)),
                              _dp(np.ascontiguousarray(goal_pos, dtype=np.float64)),
                              _dp(np.ascontiguousarray(goal_quat, dtype=np.float64)),
                              _dp(np.ascontiguousarray(q_prev_mask, dtype=np.float64)), _dp(res))
        return res

    def ik_jacobian(self, arm, q, goal_quat):
        n = self.task.arm_nmask[arm]
        jac = np.zeros((6 + 2 * n, n))
        self.L.ko_ik_jacobian(self.pm.ref(), self.d, C.byref(self.task), arm, _dp(np.ascontiguousarray(q, dtype=np.float64)),
                              _dp(np.ascontiguousarray(goal_quat, dtype=np.float64)), _dp(jac))
        return jac

    def ik_dls(self, arm, goal_pos, goal_quat, q_prev_full):
        n = self.task.arm_nmask[arm]
        out = np.zeros(n)
        self.L.ko_ik_solve_dls(self.pm.ref(), self.d, C.byref(self.task), arm,
                               _dp(np.ascontiguousarray(goal_pos, dtype=np.float64)),
                               _dp(np.ascontiguousarray(goal_quat, dtype=np.float64)),
                               _dp(np.ascontiguousarray(q_prev_full, dtype=np.float64)), _dp(out))
        return out

    def ik_trf(self, arm, goal_pos, goal_quat, q_prev_full):
        """reference ik() (ik_mujoco.py:100-155) with the genuine scipy TRF."""
        from scipy.optimize import least_squares
        mask = np.array([self.task.arm_mask[arm][i] for i in range(self.task.arm_nmask[arm])])
        rng = np.array(self.flat["jnt_range"])[mask]
        q = self.get_state()["qpos"][mask].copy()
        qprev = np.asarray(q_prev_full)[mask]
        try:
            res = least_squares(lambda x: self.ik_residual(arm, x, goal_pos, goal_quat, qprev), q,
                                jac=lambda x: self.ik_jacobian(arm, x, goal_quat), bounds=(rng[:, 0], rng[:, 1]), verbose=0)
            q = res.x
            self.ik_nfev.append(res.nfev)
        except ValueError:
            pass
        return np.clip(q, rng[:, 0], rng[:, 1])

    def _ik_trf(self, arm, gp, gq, qprev, qout):
        goal_pos = np.array([gp[i] for i in range(3)])
        goal_quat = np.array([gq[i] for i in range(4)])
        q_prev = np.array([qprev[i] for i in range(self.nq)])
        q = self.ik_trf(arm, goal_pos, goal_quat, q_prev)
        for i, v in enumerate(q):
            qout[i] = v

    # ---- flop counting
    def flops_reset(self):
        self.L.ko_flops_reset()

    def flops(self) -> int:
        return int(self.L.ko_flops_get())


def batch_step(orc: Oracle, state: Dict[str, np.ndarray], action: np.ndarray, autoreset: bool = True, seed: int = 0,
               env0: int = 0, nthreads: int = 0, ncon_peak: Optional[np.ndarray] = None):
    """Advance a batch of envs one env-step on the CPU (in place on `state`); returns obs, final_obs, reward,
    truncated, flags, ncon, contact geom pairs.  ncon_peak (optional int32 [n]) receives the most contacts any of the
    step's collision passes saw (a test diagnostic: was the cube touching anything at any sub-step)."""
    n = state["qpos"].shape[0]
    L = orc.L
    od = orc.obs_dim
    mc = L.ko_max_contacts()
    obs, fobs, rew = np.zeros((n, od)), np.zeros((n, od)), np.zeros(n)
    trunc = np.zeros(n, dtype=np.uint8)
    flags, ncon = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    geoms = np.zeros((n, 2 * mc), dtype=np.int32)
    a = np.ascontiguousarray(action, dtype=np.float32)
    ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int))  # noqa: E731
    L.ko_batch_step(orc.pm.ref(), C.byref(orc.task), C.c_int(n), _dp(state["qpos"]), _dp(state["qvel"]), _dp(state["ctrl"]),
                    _dp(state["warm"]), _dp(state["time"]), ip(state["step"]), ip(state["episode"]), _dp(state["mocap"]),
                    a.ctypes.data_as(C.POINTER(C.c_float)), _dp(obs), _dp(fobs), _dp(rew),
                    trunc.ctypes.data_as(C.POINTER(C.c_ubyte)), ip(flags), ip(ncon), ip(geoms), C.c_int(int(autoreset)),
                    C.c_ulonglong(seed), C.c_ulonglong(env0), C.c_int(nthreads), None if ncon_peak is None else ip(ncon_peak))
    return obs, fobs, rew, trunc, flags, ncon, geoms


def batch_reset_state(orc: Oracle, n: int, seed: int = 0, env0: int = 0) -> Dict[str, np.ndarray]:
    """Initial batch state: home pose + Philox cube spawn for episode 0."""
    flat, t = orc.flat, orc.task
    st = dict(qpos=np.tile(np.array(flat["qpos0"]), (n, 1)), qvel=np.zeros((n, orc.nv)), ctrl=np.zeros((n, orc.nu)),
              warm=np.zeros((n, orc.nv)), time=np.zeros(n), step=np.zeros(n, dtype=np.int32),
              episode=np.zeros(n, dtype=np.int32), mocap=np.zeros((n, 7 * max(orc.nmocap, 1))))
    home = np.array([t.q_home[i] for i in range(t.q_len)])
    st["qpos"][:, : t.q_len] = home
    st["ctrl"][:, : t.q_len] = home
    for e in range(n):
        st["qpos"][e, t.cube_qposadr: t.cube_qposadr + 3] = orc.spawn(seed, env0 + e, 0)
    mp = np.concatenate([np.concatenate([p, q]) for p, q in zip(flat["mocap_pos0"], flat["mocap_quat0"])]) if orc.nmocap else np.zeros(7)
    st["mocap"][:] = mp
    return st
