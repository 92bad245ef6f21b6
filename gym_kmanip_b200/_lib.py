"""Loader of the CUDA library (csrc/ -> lib/libkmanip_b200.so) through its C-ABI (include/kmanip_b200.h).

There is no CPU path: if the library is missing, or no CUDA device is usable, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KMANIP_B200_LIB") or os.path.join(_HERE, "lib", "libkmanip_b200.so")   # env override: A/B experiments
CSRC = os.path.join(_HERE, "csrc")
_LIB = None

EXPORTS = [
    "km_last_error", "km_version", "km_measure_fma_peak", "km_create", "km_destroy", "km_nq", "km_nv", "km_nu", "km_nmocap", "km_obs_dim",
    "km_act_dim", "km_state_dim", "km_max_contacts", "km_num_envs", "km_dtype", "km_configure", "km_set_env_ordering", "km_launch_count",
    "km_launch_config", "km_reset", "km_step", "km_get_state", "km_set_state", "km_state_ptr", "km_contacts",
    "km_solver_stats", "km_episode_stats", "km_set_host_stream", "km_debug_phase_clocks", "km_site_poses", "km_n_arm", "km_reset_host", "km_step_host",
    "km_render", "km_render_host", "km_render_record_floats", "km_get_render_records",
]


class StepOut(C.Structure):
    """km_step_out of include/kmanip_b200.h"""
    _fields_ = [("obs", C.c_void_p), ("final_obs", C.c_void_p), ("reward", C.c_void_p), ("truncated", C.c_void_p),
                ("terminated", C.c_void_p), ("con_flags", C.c_void_p), ("ncon", C.c_void_p), ("con_geoms", C.c_void_p),
                ("is_success", C.c_void_p), ("episode_return", C.c_void_p), ("final_return", C.c_void_p), ("sim_time", C.c_void_p),
                ("step_count", C.c_void_p), ("episode", C.c_void_p)]


def build(jobs: int = 8, verbose: bool = False) -> str:
    """Compile every CUDA translation unit for sm_100a into lib/libkmanip_b200.so (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, f"-j{jobs}"], stdout=out)
    return LIB_PATH


def load() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"gym_kmanip_b200: CUDA library not built ({LIB_PATH} is missing). Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C gym_kmanip_b200/csrc`. "
            "This package has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ip, u64 = C.c_void_p, C.c_int, C.c_uint64
    L.km_last_error.restype = C.c_char_p
    L.km_version.restype = C.c_char_p
    L.km_measure_fma_peak.argtypes = [ip, ip, C.POINTER(C.c_double)]
    L.km_create.argtypes = [vp, vp, ip, ip, ip, ip, u64, u64, C.POINTER(vp)]
    L.km_destroy.argtypes = [vp]
    L.km_destroy.restype = None
    for n in ("km_nq", "km_nv", "km_nu", "km_nmocap", "km_obs_dim", "km_act_dim", "km_state_dim", "km_max_contacts",
              "km_num_envs", "km_dtype"):
        getattr(L, n).argtypes = [vp]
    L.km_configure.argtypes = [vp, ip, ip]
    L.km_set_env_ordering.argtypes = [vp, ip]
    L.km_launch_count.argtypes = [vp]
    L.km_launch_count.restype = C.c_longlong
    L.km_launch_config.argtypes = [vp] + [C.POINTER(ip)] * 5
    L.km_reset.argtypes = [vp, vp, vp, vp, vp]
    L.km_step.argtypes = [vp, vp, C.POINTER(StepOut), ip, vp]
    L.km_get_state.argtypes = [vp, vp, vp, vp, vp]
    L.km_set_state.argtypes = [vp, vp, vp, vp, vp]
    L.km_state_ptr.argtypes = [vp]
    L.km_state_ptr.restype = vp
    L.km_contacts.argtypes = [vp, vp, vp, vp]
    L.km_solver_stats.argtypes = [vp, vp, vp, vp]
    L.km_debug_phase_clocks.argtypes = [vp, vp]
    L.km_episode_stats.argtypes = [vp, vp, vp, ip, vp]
    L.km_set_host_stream.argtypes = [vp, vp]
    L.km_site_poses.argtypes = [vp, vp, vp, vp]
    L.km_n_arm.argtypes = [vp]
    L.km_reset_host.argtypes = [vp, vp, vp, vp]
    L.km_step_host.argtypes = [vp, vp, vp, vp, vp, ip]
    L.km_render.argtypes = [vp, vp, vp, vp, vp]
    L.km_render_host.argtypes = [vp, vp, vp, vp]
    L.km_render_record_floats.argtypes = [vp]
    L.km_get_render_records.argtypes = [vp, vp, vp]
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"kmanip_b200 error {rc}: {load().km_last_error().decode()}")
