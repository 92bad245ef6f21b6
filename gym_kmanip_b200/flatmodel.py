"""ctypes views of the flat model / task structs that cross the C-ABI (include/kmanip_b200.h).

The same plain-C layout is accepted by the CUDA library (km_model / km_task) and by the test
oracle (oracle/ko_model.h); this module only packs numbers, it computes nothing.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np

from . import constants as K

MAXARM, MAXMASK = 2, 8

_INT_FIELDS_HEAD = ["nbody", "njnt", "nq", "nv", "nu", "nsite", "ngeom", "npair", "nmocap", "iterations", "ls_iterations"]
_PTRS = [  # (name, ctype, flat key)
    ("body_parent", C.c_int), ("body_rootid", C.c_int), ("body_mocapid", C.c_int), ("body_jntadr", C.c_int), ("body_jntnum", C.c_int),
    ("body_pos", C.c_double), ("body_quat", C.c_double), ("body_mass", C.c_double), ("body_ipos", C.c_double),
    ("body_inertia", C.c_double), ("body_invweight0", C.c_double),
    ("jnt_type", C.c_int), ("jnt_bodyid", C.c_int), ("jnt_qposadr", C.c_int), ("jnt_dofadr", C.c_int), ("jnt_limited", C.c_int),
    ("jnt_pos", C.c_double), ("jnt_axis", C.c_double), ("jnt_range", C.c_double), ("jnt_solref", C.c_double),
    ("jnt_solimp", C.c_double), ("qpos0", C.c_double),
    ("dof_bodyid", C.c_int), ("dof_jntid", C.c_int), ("dof_parentid", C.c_int),
    ("dof_frictionloss", C.c_double), ("dof_solref", C.c_double), ("dof_solimp", C.c_double), ("dof_invweight0", C.c_double),
    ("act_jntid", C.c_int), ("act_ctrllimited", C.c_int), ("act_forcelimited", C.c_int),
    ("act_kp", C.c_double), ("act_ctrlrange", C.c_double), ("act_forcerange", C.c_double),
    ("site_bodyid", C.c_int), ("site_pos", C.c_double), ("site_quat", C.c_double),
    ("geom_type", C.c_int), ("geom_bodyid", C.c_int), ("geom_pos", C.c_double), ("geom_quat", C.c_double), ("geom_size", C.c_double),
    ("pair_geom1", C.c_int), ("pair_geom2", C.c_int), ("pair_condim", C.c_int),
    ("pair_friction", C.c_double), ("pair_solref", C.c_double), ("pair_solimp", C.c_double), ("pair_margin", C.c_double),
    ("mocap_pos0", C.c_double), ("mocap_quat0", C.c_double),
]


class CModel(C.Structure):
    _fields_ = ([(n, C.c_int) for n in _INT_FIELDS_HEAD]
                + [("timestep", C.c_double), ("gravity", C.c_double * 3), ("tolerance", C.c_double),
                   ("ls_tolerance", C.c_double), ("impratio", C.c_double), ("meaninertia", C.c_double)]
                + [(n, C.POINTER(t)) for n, t in _PTRS])


class CTask(C.Structure):
    _fields_ = [
        ("q_len", C.c_int), ("n_arm", C.c_int), ("act_dim", C.c_int), ("act_mode", C.c_int),
        ("arm_nmask", C.c_int * MAXARM), ("arm_mask", (C.c_int * MAXMASK) * MAXARM), ("arm_grip", (C.c_int * 2) * MAXARM),
        ("arm_site", C.c_int * MAXARM), ("arm_eebody", C.c_int * MAXARM), ("arm_mocap", C.c_int * MAXARM),
        ("off_pos", C.c_int * MAXARM), ("off_orn", C.c_int * MAXARM), ("off_grip", C.c_int * MAXARM), ("off_q", C.c_int * MAXARM),
        ("cube_body", C.c_int), ("cube_qposadr", C.c_int), ("ik_iters", C.c_int), ("ik_teleport", C.c_int),
        ("max_episode_steps", C.c_int), ("ik_mode", C.c_int), ("n_sub_steps", C.c_int), ("reserved0", C.c_int),
        ("q_home", C.c_double * 32), ("cube_spawn_lo", C.c_double * 3), ("cube_spawn_hi", C.c_double * 3),
    ]


class PackedModel:
    """Owns the numpy buffers behind a CModel."""

    def __init__(self, flat: Dict):
        self.flat = flat
        self.c = CModel()
        self._keep: List[np.ndarray] = []
        for n in _INT_FIELDS_HEAD[:9]:
            setattr(self.c, n, int(flat[n]))
        opt = flat["opt"]
        self.c.iterations = int(opt["iterations"])
        self.c.ls_iterations = int(opt["ls_iterations"])
        self.c.timestep = opt["timestep"]
        self.c.gravity = (C.c_double * 3)(*opt["gravity"])
        self.c.tolerance = opt["tolerance"]
        self.c.ls_tolerance = opt["ls_tolerance"]
        self.c.impratio = opt["impratio"]
        self.c.meaninertia = flat["meaninertia"]
        for name, ct in _PTRS:
            arr = np.ascontiguousarray(np.array(flat[name], dtype=np.int32 if ct is C.c_int else np.float64).reshape(-1))
            if arr.size == 0:
                arr = np.zeros(1, dtype=arr.dtype)
            self._keep.append(arr)
            setattr(self.c, name, arr.ctypes.data_as(C.POINTER(ct)))

    def ref(self):
        return C.byref(self.c)


def action_layout(act_list: List[str], n_r: int, n_l: int) -> Dict[str, slice]:
    """Flat action layout: present keys in the order env_base.py:149-190 builds the action Dict."""
    sizes = dict(eel_pos=3, eel_orn=3, eer_pos=3, eer_orn=3, grip_l=1, grip_r=1, q_pos_r=n_r, q_pos_l=n_l)
    out, o = {}, 0
    for k in K.ACTION_KEY_ORDER:
        if k in act_list:
            out[k] = slice(o, o + sizes[k])
            o += sizes[k]
    return out


def make_task(flat: Dict, env_kwargs: Dict, ik_iters: int = K.DEVICE_IK_ITERS, ik_teleport: bool = True,
              max_episode_steps: int = K.MAX_EPISODE_STEPS, ik_mode: int = 0, n_sub_steps: int = 0) -> CTask:
    """Task struct for one registered env id (kwargs as in constants.ENV_REGISTRY)."""
    act_list = env_kwargs["act_list"]
    t = CTask()
    home = np.asarray(env_kwargs["q_pos_home"], dtype=np.float32)
    t.q_len = len(home)
    for i, v in enumerate(home):
        t.q_home[i] = float(v)
    masks = [env_kwargs.get("q_id_r_mask"), env_kwargs.get("q_id_l_mask")]
    grips = [env_kwargs.get("ctrl_id_r_grip"), env_kwargs.get("ctrl_id_l_grip")]
    sides = ["r", "l"]
    n_r = len(masks[0]) if masks[0] is not None else 0
    n_l = len(masks[1]) if masks[1] is not None else 0
    lay = action_layout(act_list, n_r, n_l)
    t.act_dim = max([s.stop for s in lay.values()] + [0])
    t.act_mode = 1 if any(k.startswith("q_pos_") for k in act_list) else 0
    assert not (t.act_mode == 1 and any(k.endswith("_pos") and k.startswith("ee") for k in act_list)), \
        "mixed end-effector and joint-position actions are not a registered configuration"
    n_arm = 0
    for a, s in enumerate(sides):
        present = any(k in act_list for k in (f"ee{s}_pos", f"grip_{s}", f"q_pos_{s}"))
        if not present:
            continue
        assert a == n_arm, "a left-only action list is not a registered configuration"
        n_arm += 1
        t.arm_nmask[a] = len(masks[a])
        for i, j in enumerate(masks[a]):
            t.arm_mask[a][i] = int(j)
        t.arm_grip[a][0], t.arm_grip[a][1] = int(grips[a][0]), int(grips[a][1])
        t.arm_site[a] = flat["site_name"].index(f"ee{s}_site_pos")
        t.arm_eebody[a] = flat["body_name"].index(f"ee{s}_site")
        t.arm_mocap[a] = K.MOCAP_ID_R if s == "r" else K.MOCAP_ID_L
        t.off_pos[a] = lay[f"ee{s}_pos"].start if f"ee{s}_pos" in lay else -1
        t.off_orn[a] = lay[f"ee{s}_orn"].start if f"ee{s}_orn" in lay else -1
        t.off_grip[a] = lay[f"grip_{s}"].start if f"grip_{s}" in lay else -1
        t.off_q[a] = lay[f"q_pos_{s}"].start if f"q_pos_{s}" in lay else -1
    t.n_arm = n_arm
    t.cube_body = flat["body_name"].index("cube")
    t.cube_qposadr = flat["jnt_qposadr"][flat["jnt_name"].index("cube_joint")]
    t.ik_iters = ik_iters
    t.ik_teleport = int(ik_teleport)
    t.max_episode_steps = max_episode_steps
    t.ik_mode = int(ik_mode)
    t.n_sub_steps = int(n_sub_steps)    # 0: the reference's 10 (control timestep 0.02 s / physics timestep 0.002 s)
    for i in range(3):
        t.cube_spawn_lo[i] = float(K.CUBE_SPAWN_RANGE[i, 0])
        t.cube_spawn_hi[i] = float(K.CUBE_SPAWN_RANGE[i, 1])
    return t


def obs_layout(q_len: int) -> Dict[str, slice]:
    """Flat observation layout of the state obs (env_sim.py:110-139)."""
    return dict(q_pos=slice(0, q_len), q_vel=slice(q_len, 2 * q_len), cube_pos=slice(2 * q_len, 2 * q_len + 3),
                cube_orn=slice(2 * q_len + 3, 2 * q_len + 7))
