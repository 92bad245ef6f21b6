"""ctypes front-end of tests/hostsim/hostsim.cpp (the kernel body built for the host, one lane per env).

TEST INFRASTRUCTURE ONLY -- see hostsim.cpp.  Used to check the simulator source stage by stage against the
oracle on machines without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _ROOT)
from gym_kmanip_b200 import constants as K, flatmodel, mjcf   # noqa: E402

_SO = os.path.join(_HERE, "_build", "libhostsim.so")
_SRC = [os.path.join(_HERE, "hostsim.cpp")] + [os.path.join(_ROOT, "gym_kmanip_b200", "csrc", f) for f in
                                                ("km_common.cuh", "km_model.cuh", "km_sim.cuh", "km_solver_tpe.cuh", "km_solver_warp.cuh", "km_ik_trf.cuh", "km_fill.h")] + \
    [os.path.join(_HERE, "warpemu.h")]
_LIB = None
SCENE_ID = {"solo_arm": 0, "dual_arm": 1, "torso": 2}


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in _SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-g1", "-fPIC", "-std=c++17", "-x", "c++", "-shared", "-ffp-contract=off", "-o", _SO,
                               _SRC[0], "-ldl"])


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(_SO)
        L.hs_create.restype = C.c_void_p
        L.hs_last_error.restype = C.c_char_p
        L.hs_fault.restype = C.c_char_p
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class HostSim:
    def __init__(self, env_id="KManipSoloArm", dtype=64, lanes=1, **task_kw):
        """lanes = 1: generic code, one lane per env; lanes = 32: the warp-per-env device code on an emulated warp."""
        self.kw = K.ENV_REGISTRY[env_id]
        self.scene = mjcf.scene_of_mjcf(self.kw["mjcf_filename"])
        self.flat = mjcf.load_flat(self.scene)
        self.pm = flatmodel.PackedModel(self.flat)
        self.task = flatmodel.make_task(self.flat, self.kw, **task_kw)
        self.L = lib()
        h = self.L.hs_create(self.pm.ref(), C.byref(self.task), SCENE_ID[self.scene], dtype, lanes)
        if not h:
            raise RuntimeError(self.L.hs_last_error().decode())
        self.h = C.c_void_p(h)
        d = (C.c_int * 8)()
        self.L.hs_dims(self.h, d)
        self.nq, self.nv, self.nu, self.nmocap, self.obs_dim, self.act_dim, self.maxcon = list(d)[:7]
        self.state_dim = self.nq + 2 * self.nv + self.nu + 7 * self.nmocap + 1 + 3

    def __del__(self):
        try:
            self.L.hs_destroy(self.h)
        except Exception:
            pass

    def pack(self, st):
        return np.concatenate([st["qpos"], st["qvel"], st["ctrl"], st["warm"], st["mocap"], [st["time"]],
                               st.get("cube_lo", np.zeros(3))]).astype(np.float64)

    def unpack(self, rec):
        o = 0
        out = {}
        for k, n in (("qpos", self.nq), ("qvel", self.nv), ("ctrl", self.nu), ("warm", self.nv), ("mocap", 7 * self.nmocap)):
            out[k] = rec[o:o + n].copy()
            o += n
        out["time"] = float(rec[o])
        out["cube_lo"] = rec[o + 1:o + 4].copy()
        out["qpos"][-7:-4] += out["cube_lo"]      # the cube's position is hi + lo (fp32 build; lo = 0 in fp64)
        return out

    def set_state(self, st, step=0, episode=0):
        rec = np.ascontiguousarray(self.pack(st))
        self.L.hs_set_state(self.h, _dp(rec), C.c_int(step), C.c_int(episode))

    def get_state(self):
        rec = np.zeros(self.state_dim)
        s, e = C.c_int(0), C.c_int(0)
        self.L.hs_get_state(self.h, _dp(rec), C.byref(s), C.byref(e))
        out = self.unpack(rec)
        out["step"], out["episode"] = s.value, e.value
        return out

    def fault(self):
        """Non-empty when the lanes of the emulated warp diverged at a collective (a hang on the GPU)."""
        return self.L.hs_fault(self.h).decode()

    def step1(self):
        self.L.hs_step1(self.h)

    def step2(self):
        self.L.hs_step2(self.h)

    def before_step(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        self.L.hs_before_step(self.h, a.ctypes.data_as(C.POINTER(C.c_float)))

    def env_step(self, a, autoreset=False, seed=0, env_id=0):
        a = np.ascontiguousarray(a, dtype=np.float32)
        obs, fobs = np.zeros(self.obs_dim), np.zeros(self.obs_dim)
        r = C.c_double(0)
        tr = C.c_ubyte(0)
        fl, nc = C.c_int(0), C.c_int(0)
        geoms = np.zeros(2 * self.maxcon, dtype=np.int32)
        self.L.hs_env_step(self.h, a.ctypes.data_as(C.POINTER(C.c_float)), _dp(obs), _dp(fobs), C.byref(r), C.byref(tr),
                           C.byref(fl), C.byref(nc), geoms.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(int(autoreset)),
                           C.c_ulonglong(seed), C.c_ulonglong(env_id))
        return dict(obs=obs, final_obs=fobs, reward=r.value, truncated=bool(tr.value), flags=fl.value, ncon=nc.value, geoms=geoms)

    def reset(self, seed=0, env_id=0, xyz=None):
        obs = np.zeros(self.obs_dim)
        p = None if xyz is None else _dp(np.ascontiguousarray(xyz, dtype=np.float64))
        self.L.hs_reset(self.h, C.c_ulonglong(seed), C.c_ulonglong(env_id), p, _dp(obs))
        return obs

    def field(self, name, cap=65536):
        out = np.zeros(cap)
        n = self.L.hs_field(self.h, name.encode(), _dp(out), cap)
        if n < 0:
            raise KeyError(name)
        return out[:n].copy()
