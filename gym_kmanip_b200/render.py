"""Host side of the camera observations: MJCF cameras / lights of the flat model -> km_camera / km_visual (include/kmanip_b200.h).

The device works with moving links only (static bodies are folded away, csrc/km_fill.h), so a camera fixed to a body is
re-expressed in the frame of the closest jointed ancestor body (or the world), exactly as the flattener's kinematics
would place it.  Camera sizes come from the reference's Cam table (reference __init__.py:157-161, constants.CAMERAS).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np

from . import mjcf

V3 = C.c_double * 3


class CCamera(C.Structure):
    """km_camera"""
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fovy", C.c_double), ("link", C.c_int), ("pos", V3),
                ("target_link", C.c_int), ("target_pos", V3)]


class CVisual(C.Structure):
    """km_visual"""
    _fields_ = [("nlight", C.c_int), ("light_dir", V3 * 4), ("light_diffuse", V3 * 4), ("light_specular", V3 * 4),
                ("light_ambient", V3 * 4), ("head_ambient", V3), ("head_diffuse", V3), ("head_specular", V3),
                ("rgb_table", V3), ("rgb_cube", V3), ("rgb_link", V3), ("rgb_pad", V3), ("specular", C.c_double),
                ("shininess", C.c_double), ("link_radius", C.c_double)]


def point_in_link(flat: Dict, body: int, pos) -> Tuple[int, np.ndarray]:
    """A point given in `body`'s frame -> (joint index of the closest jointed ancestor-or-self, point in its frame);
    joint index -1 means the world frame (the body does not move)."""
    p = np.asarray(pos, dtype=np.float64)
    b = int(body)
    while b != 0 and flat["body_jntnum"][b] == 0:
        p = np.asarray(flat["body_pos"][b]) + mjcf.quat_to_mat(np.asarray(flat["body_quat"][b])) @ p
        b = flat["body_parent"][b]
    if b == 0:
        return -1, p
    j = flat["body_jntadr"][b]
    assert flat["jnt_type"][j] in (mjcf.JNT_SLIDE, mjcf.JNT_HINGE), "cameras ride on articulated links"
    return j, p


def camera_struct(flat: Dict, cam_name: str, width: int, height: int) -> CCamera:
    if cam_name not in flat.get("cam_name", []):
        raise KeyError(f"camera {cam_name!r} is not part of scene {flat['scene']} (cameras: {flat.get('cam_name')})")
    ci = flat["cam_name"].index(cam_name)
    link, pos = point_in_link(flat, flat["cam_bodyid"][ci], flat["cam_pos"][ci])
    tlink, tpos = point_in_link(flat, flat["cam_targetbodyid"][ci], [0.0, 0.0, 0.0])
    return CCamera(int(width), int(height), float(flat["cam_fovy"][ci]), link, V3(*pos), tlink, V3(*tpos))


def visual_struct(flat: Dict) -> CVisual:
    v = CVisual()
    n = len(flat["light_dir"])
    assert n <= 4
    v.nlight = n
    for l in range(n):
        v.light_dir[l] = V3(*flat["light_dir"][l])
        v.light_diffuse[l] = V3(*flat["light_diffuse"][l])
        v.light_specular[l] = V3(*flat["light_specular"][l])
        v.light_ambient[l] = V3(*flat["light_ambient"][l])
    h = flat["headlight"]
    v.head_ambient, v.head_diffuse, v.head_specular = V3(*h["ambient"]), V3(*h["diffuse"]), V3(*h["specular"])
    gt, gn = flat["geom_type"], flat["geom_name"]
    v.rgb_table = V3(*flat["geom_rgba"][gt.index(mjcf.GEOM_PLANE)][:3])
    v.rgb_cube = V3(*flat["geom_rgba"][gn.index("cube")][:3])
    v.rgb_pad = V3(*flat["geom_rgba"][gt.index(mjcf.GEOM_SPHERE)][:3])
    vis = flat["visual"]
    v.rgb_link = V3(*vis["link_rgba"][:3])
    v.specular, v.shininess, v.link_radius = vis["specular"], vis["shininess"], vis["link_radius"]
    return v
