"""Constants and env registry of the drop-in boundary.

These restate the API contract of reference gym_kmanip/__init__.py:28-208 (time steps, deltas,
reward weights, home poses, index masks) and the kwargs of its eight ``register()`` calls
(reference gym_kmanip/__init__.py:244-483).  Values only -- no reference code is reused.
"""
from __future__ import annotations

import os as _os
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

SOLO_ARM_MJCF = "_env_solo_arm.xml"
DUAL_ARM_MJCF = "_env_dual_arm.xml"
TORSO_MJCF = "_env_torso.xml"
SOLO_ARM_URDF = "stompy_tiny_solo_arm_glb.urdf"
DUAL_ARM_URDF = "stompy_dual_arm_tiny_glb.urdf"
TORSO_URDF = "stompy_tiny_glb/robot.urdf"

DATA_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "data")   # reference __init__.py:12
DATE_FORMAT = "%mm%dd%Yy_%Hh%Mm"   # :15
MAX_EPISODE_STEPS = 64            # reference __init__.py:28
FPS = 30                          # :29
CONTROL_TIMESTEP = 0.02           # :30
MAX_Q_VEL = float(np.pi)          # :31
CTRL_ALPHA = 1.0                  # :34
IK_RES_RAD, IK_RES_REG_PREV, IK_RES_REG_HOME = 0.02, 6e-3, 2e-6   # :37-39
IK_JAC_RAD, IK_JAC_REG = 0.02, 9e-3                                 # :40-41
OBS_DTYPE = np.float64            # :50
ACT_DTYPE = np.float32            # :51

MOCAP_ID_R, MOCAP_ID_L = 0, 1     # :139-140
CUBE_SPAWN_RANGE = np.array([[0.1, 0.3], [0.5, 0.7], [0.6, 0.7]])   # :164-170
EE_POS_DELTA = np.array([0.01, 0.01, 0.01])    # :174-180
EE_ORN_DELTA = np.array([0.1, 0.1, 0.1])       # :181-187
EPSILON = 1e-6                    # :190
Q_POS_DELTA = 0.1                 # :196
EE_S_MIN, EE_S_MAX, EE_S_DELTA = -0.029, 0.005, 0.0001   # :199-201
REWARD_SUCCESS_THRESHOLD = 2.0    # :204
REWARD_VEL_PENALTY, REWARD_GRIP_DIST = 0.01, 0.01        # :205-206
REWARD_TOUCH_CUBE, REWARD_LIFT_CUBE = 1.0, 1.0           # :207-208
XYZW_2_WXYZ = np.array([3, 0, 1, 2])
WXYZ_2_XYZW = np.array([1, 2, 3, 0])

DEVICE_IK_ITERS = 6               # damped Gauss-Newton iterations of the batched IK


def _home(pairs: List[Tuple[str, float]]):
    d = OrderedDict(pairs)
    return d, np.array(list(d.values()), dtype=ACT_DTYPE), list(d.keys())


_R = "joint_right_arm_1_"
_L1 = "joint_left_arm_1_"   # the reference's dual-arm dict uses *_arm_1_* names for the left arm
_L2 = "joint_left_arm_2_"

Q_SOLO_ARM_HOME_DICT, Q_SOLO_ARM_HOME, Q_SOLO_ARM_KEYS = _home([
    (_R + "x8_1_dof_x8", 0.0), (_R + "x8_2_dof_x8", 0.75), (_R + "x6_1_dof_x6", 1.0), (_R + "x6_2_dof_x6", 1.0),
    (_R + "x4_1_dof_x4", 2.0), (_R + "hand_right_1_x4_3_dof_x4", -2.0), (_R + "hand_right_1_x4_1_dof_x4", 0.0),
    (_R + "hand_right_1_x4_2_dof_x4", 0.0), (_R + "hand_right_1_slider_3", 0.005), (_R + "hand_right_1_slider_1", 0.005),
])   # reference __init__.py:53-67
Q_DUAL_ARM_HOME_DICT, Q_DUAL_ARM_HOME, Q_DUAL_ARM_KEYS = _home([
    (_R + "x8_1_dof_x8", 0.0), (_R + "x8_2_dof_x8", 0.75), (_R + "x6_1_dof_x6", 1.0), (_R + "x6_2_dof_x6", 1.0),
    (_R + "x4_1_dof_x4", 2.0), (_R + "hand_right_1_x4_3_dof_x4", -2.7), (_R + "hand_right_1_x4_1_dof_x4", 0.0),
    (_R + "hand_right_1_x4_2_dof_x4", 0.0), (_R + "hand_right_1_slider_3", 0.005), (_R + "hand_right_1_slider_1", 0.005),
    (_L1 + "x8_1_dof_x8", 0.0), (_L1 + "x8_2_dof_x8", -0.75), (_L1 + "x6_1_dof_x6", -1.0), (_L1 + "x6_2_dof_x6", -1.0),
    (_L1 + "x4_1_dof_x4", 2.0), (_L1 + "hand_left_1_x4_3_dof_x4", 0.0), (_L1 + "hand_left_1_x4_1_dof_x4", 0.0),
    (_L1 + "hand_left_1_x4_2_dof_x4", 0.0), (_L1 + "hand_left_1_slider_3", 0.005), (_L1 + "hand_left_1_slider_1", 0.005),
])   # reference __init__.py:69-94
Q_TORSO_HOME_DICT, Q_TORSO_HOME, Q_TORSO_KEYS = _home([
    ("joint_head_1_x4_1_dof_x4", -1.0), ("joint_head_1_x4_2_dof_x4", 0.0),
    (_R + "x8_1_dof_x8", 1.7), (_R + "x8_2_dof_x8", 1.6), (_R + "x6_1_dof_x6", 0.34), (_R + "x6_2_dof_x6", 1.6),
    (_R + "x4_1_dof_x4", 1.4), (_R + "hand_1_x4_1_dof_x4", -0.26), (_R + "hand_1_slider_1", 0.0),
    (_R + "hand_1_slider_2", 0.0), (_R + "hand_1_x4_2_dof_x4", 0.0),
    (_L2 + "x8_1_dof_x8", -1.7), (_L2 + "x8_2_dof_x8", -1.6), (_L2 + "x6_1_dof_x6", -0.34), (_L2 + "x6_2_dof_x6", -1.6),
    (_L2 + "x4_1_dof_x4", -1.4), (_L2 + "hand_1_x4_1_dof_x4", -1.7), (_L2 + "hand_1_slider_1", 0.0),
    (_L2 + "hand_1_slider_2", 0.0), (_L2 + "hand_1_x4_2_dof_x4", 0.0),
])   # reference __init__.py:96-121

Q_ID_R_MASK_SOLO = np.array([0, 1, 2, 3, 4, 5, 6]);  CTRL_ID_R_GRIP_SOLO = np.array([8, 9])            # :125-126
Q_ID_R_MASK_DUAL = np.array([0, 1, 2, 3, 4, 5, 6]);  Q_ID_L_MASK_DUAL = np.array([10, 11, 12, 13, 14, 15, 16])
CTRL_ID_R_GRIP_DUAL = np.array([8, 9]);              CTRL_ID_L_GRIP_DUAL = np.array([18, 19])          # :128-131
Q_ID_R_MASK_TORSO = np.array([2, 3, 4, 5, 6, 7]);    Q_ID_L_MASK_TORSO = np.array([11, 12, 13, 14, 15, 16])
CTRL_ID_R_GRIP_TORSO = np.array([8, 9]);             CTRL_ID_L_GRIP_TORSO = np.array([17, 18])         # :133-136


@dataclass
class Cam:                       # reference __init__.py:143-154
    w: int
    h: int
    c: int
    fl: int
    pp: Tuple[int, int]
    name: str
    log_name: str
    low: int = 0
    high: int = 255
    dtype = np.uint8


CAMERAS: "OrderedDict[str, Cam]" = OrderedDict(
    head=Cam(640, 480, 3, 448, (320, 240), "head", "camera/head"),
    top=Cam(640, 480, 3, 448, (320, 240), "top", "camera/top"),
    grip_r=Cam(60, 40, 3, 45, (30, 20), "grip_r", "camera/grip_r"),
    grip_l=Cam(60, 40, 3, 45, (30, 20), "grip_l", "camera/grip_l"),
)

_STATE_OBS = ["q_pos", "q_vel", "cube_pos", "cube_orn"]
_EE_SOLO = ["eer_pos", "eer_orn", "grip_r"]
_EE_DUAL = ["eel_pos", "eel_orn", "eer_pos", "eer_orn", "grip_l", "grip_r"]
_SOLO = dict(mjcf_filename=SOLO_ARM_MJCF, urdf_filename=SOLO_ARM_URDF, q_pos_home=Q_SOLO_ARM_HOME,
             q_dict=Q_SOLO_ARM_HOME_DICT, q_keys=Q_SOLO_ARM_KEYS, q_id_r_mask=Q_ID_R_MASK_SOLO,
             ctrl_id_r_grip=CTRL_ID_R_GRIP_SOLO)
_DUAL = dict(mjcf_filename=DUAL_ARM_MJCF, urdf_filename=DUAL_ARM_URDF, q_pos_home=Q_DUAL_ARM_HOME,
             q_dict=Q_DUAL_ARM_HOME_DICT, q_keys=Q_DUAL_ARM_KEYS, q_id_r_mask=Q_ID_R_MASK_DUAL,
             q_id_l_mask=Q_ID_L_MASK_DUAL, ctrl_id_r_grip=CTRL_ID_R_GRIP_DUAL, ctrl_id_l_grip=CTRL_ID_L_GRIP_DUAL)
_TORSO = dict(mjcf_filename=TORSO_MJCF, urdf_filename=TORSO_URDF, q_pos_home=Q_TORSO_HOME,
              q_dict=Q_TORSO_HOME_DICT, q_keys=Q_TORSO_KEYS, q_id_r_mask=Q_ID_R_MASK_TORSO,
              q_id_l_mask=Q_ID_L_MASK_TORSO, ctrl_id_r_grip=CTRL_ID_R_GRIP_TORSO, ctrl_id_l_grip=CTRL_ID_L_GRIP_TORSO)

# id -> constructor kwargs, as registered by the reference (__init__.py:244-483)
ENV_REGISTRY: Dict[str, dict] = {
    "KManipSoloArm": dict(_SOLO, obs_list=list(_STATE_OBS), act_list=list(_EE_SOLO)),
    "KManipSoloArmQPos": dict(_SOLO, obs_list=list(_STATE_OBS), act_list=["q_pos_r", "grip_r"]),
    "KManipSoloArmVision": dict(_SOLO, obs_list=["q_pos", "q_vel", "camera/head", "camera/grip_r"], act_list=list(_EE_SOLO)),
    "KManipDualArm": dict(_DUAL, obs_list=list(_STATE_OBS), act_list=list(_EE_DUAL)),
    "KManipDualArmQPos": dict(_DUAL, obs_list=list(_STATE_OBS), act_list=["q_pos_r", "q_pos_l", "grip_l", "grip_r"]),
    "KManipDualArmVision": dict(_DUAL, obs_list=["q_pos", "q_vel", "camera/head", "camera/grip_l", "camera/grip_r"],
                                act_list=list(_EE_DUAL)),
    "KManipTorso": dict(_TORSO, obs_list=list(_STATE_OBS), act_list=list(_EE_DUAL)),
    "KManipTorsoVision": dict(_TORSO, obs_list=["q_pos", "q_vel", "camera/head", "camera/grip_l", "camera/grip_r"],
                              act_list=list(_EE_DUAL)),
}

# order in which env_base.py:149-190 inserts keys into the action Dict; the flat action vector of the
# batched API concatenates the present keys in this order
ACTION_KEY_ORDER = ["eel_pos", "eel_orn", "eer_pos", "eer_orn", "grip_l", "grip_r", "q_pos_r", "q_pos_l"]
