"""Generate the compile-time scene topology headers for the CUDA kernels.

The B200 path compiles each MJCF scene *into* its kernel: tree structure, joint kinds, constraint
slots and the task's index masks become ``static constexpr`` tables (csrc/scenes/scene_<name>.h) that the
kernel templates unroll over, so per-env quantities live in registers with static indexing.  All
numeric parameters (poses, inertias, gains, solver parameters) stay run-time values passed to the kernel
as a by-value parameter struct (constant bank), so they can be changed without recompiling.

The generator also *checks* the structural assumptions the kernels exploit (one joint per body,
hinge/slide axes = local z through the body origin, identity site/geom/inertial frames, horizontal
table plane, cube COM at its origin).
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np

from . import constants as K
from .mjcf import JNT_FREE, JNT_HINGE, JNT_SLIDE, GEOM_PLANE, GEOM_SPHERE, GEOM_BOX

_HERE = os.path.dirname(os.path.abspath(__file__))
SCENE_DIR = os.path.join(_HERE, "csrc", "scenes")

SCENE_STRUCT = {"solo_arm": "SceneSoloArm", "dual_arm": "SceneDualArm", "torso": "SceneTorso"}
SCENE_ENV = {"solo_arm": "KManipSoloArm", "dual_arm": "KManipDualArm", "torso": "KManipTorso"}
SCENE_ID = {"solo_arm": 0, "dual_arm": 1, "torso": 2}


def topology(scene: str, flat: Dict) -> Dict:
    """Derive (and validate) the structural tables of one scene."""
    nb, nv, nq, nu = flat["nbody"], flat["nv"], flat["nq"], flat["nu"]
    kw = K.ENV_REGISTRY[SCENE_ENV[scene]]
    t: Dict = dict(NBODY=nb, NQ=nq, NV=nv, NU=nu, NMOCAP=flat["nmocap"], Q_LEN=len(kw["q_pos_home"]))
    jtype, qadr, dadr = [-1] * nb, [-1] * nb, [-1] * nb
    for b in range(nb):
        assert flat["body_jntnum"][b] <= 1, "kernels assume at most one joint per body"
        if flat["body_jntnum"][b]:
            j = flat["body_jntadr"][b]
            jtype[b], qadr[b], dadr[b] = flat["jnt_type"][j], flat["jnt_qposadr"][j], flat["jnt_dofadr"][j]
            if jtype[b] in (JNT_HINGE, JNT_SLIDE):
                assert np.allclose(flat["jnt_axis"][j], [0, 0, 1]) and np.allclose(flat["jnt_pos"][j], 0)
                assert flat["qpos0"][qadr[b]] == 0
            else:
                assert jtype[b] == JNT_FREE and flat["body_parent"][b] == 0
                assert np.allclose(flat["body_ipos"][b], 0), "free body COM must be at its origin"
    moving = [0] * nb
    lastdof = [-1] * nb
    chain = [-1] * nb
    for b in range(1, nb):
        p = flat["body_parent"][b]
        moving[b] = 1 if (jtype[b] >= 0 or moving[p]) else 0
        lastdof[b] = lastdof[p]
        if jtype[b] >= 0:
            lastdof[b] = dadr[b] + (5 if jtype[b] == JNT_FREE else 0)
        if moving[b]:
            chain[b] = chain[p] if moving[p] else b   # chain id = its first moving body
    t.update(body_parent=flat["body_parent"], body_jtype=jtype, body_qadr=qadr, body_dadr=dadr, body_moving=moving,
             body_lastdof=lastdof, body_chain=chain, body_mocap=flat["body_mocapid"],
             body_hasmass=[1 if m > 0 else 0 for m in flat["body_mass"]],
             dof_parent=flat["dof_parentid"], dof_body=flat["dof_bodyid"])
    cube_b = flat["body_name"].index("cube")
    assert jtype[cube_b] == JNT_FREE and dadr[cube_b] == nv - 6 and qadr[cube_b] == nq - 7
    t.update(CUBE_BODY=cube_b, NVA=nv - 6, NQA=nq - 7)   # articulated dofs come first, the free cube last
    for b in range(1, nb):
        if jtype[b] in (JNT_HINGE, JNT_SLIDE):
            assert dadr[b] == qadr[b] < t["NVA"], "hinge/slide joints must precede the cube and map q==dof index"
    # actuators: one <position> per articulated joint, in joint order (reference relies on it: SURVEY.md B-6)
    assert nu == t["NVA"]
    for i in range(nu):
        assert flat["jnt_dofadr"][flat["act_jntid"][i]] == i
    # joint limits: every articulated joint is limited
    for j in range(flat["njnt"]):
        if flat["jnt_type"][j] != JNT_FREE:
            assert flat["jnt_limited"][j] and flat["jnt_dofadr"][j] == j
    t["fric_dof"] = [d for d in range(nv) if flat["dof_frictionloss"][d] > 0]
    t["NFRIC"] = len(t["fric_dof"])
    # geoms / pairs
    gt, gb = flat["geom_type"], flat["geom_bodyid"]
    cube_g = flat["geom_name"].index("cube")
    table_g = flat["geom_name"].index("table")
    assert gt[cube_g] == GEOM_BOX and gb[cube_g] == cube_b and np.allclose(flat["geom_pos"][cube_g], 0)
    assert np.allclose(flat["geom_quat"][cube_g], [1, 0, 0, 0])
    assert gt[table_g] == GEOM_PLANE and not moving[gb[table_g]] and np.allclose(flat["geom_quat"][table_g], [1, 0, 0, 0])
    tb = gb[table_g]
    assert flat["body_parent"][tb] == 0 and np.allclose(flat["body_quat"][tb], [1, 0, 0, 0])
    pads, pad_pair, table_pair = [], [], -1
    for p in range(flat["npair"]):
        g1, g2 = flat["pair_geom1"][p], flat["pair_geom2"][p]
        assert g2 == cube_g and flat["pair_condim"][p] == 4 and flat["pair_margin"][p] == 0
        if g1 == table_g:
            table_pair = p
        else:
            assert gt[g1] == GEOM_SPHERE
            pads.append(g1)
            pad_pair.append(p)
    assert table_pair == flat["npair"] - 1 and pad_pair == list(range(len(pads))), "pads first, then the table"
    t.update(NPAD=len(pads), pad_geom=pads, pad_body=[gb[g] for g in pads], TABLE_GEOM=table_g, CUBE_GEOM=cube_g,
             TABLE_BODY=tb)
    # task (arms in processing order right, left)
    masks = [kw.get("q_id_r_mask"), kw.get("q_id_l_mask")]
    grips = [kw.get("ctrl_id_r_grip"), kw.get("ctrl_id_l_grip")]
    narm = 2 if masks[1] is not None else 1
    t["NARM"] = narm
    t["arm_nmask"] = [len(masks[a]) for a in range(narm)]
    t["arm_mask"] = [list(map(int, masks[a])) + [-1] * (8 - len(masks[a])) for a in range(narm)]
    t["arm_grip"] = [list(map(int, grips[a])) for a in range(narm)]
    sb = []
    for a, s in enumerate(["r", "l"][:narm]):
        sid = flat["site_name"].index(f"ee{s}_site_pos")
        assert np.allclose(flat["site_pos"][sid], 0) and np.allclose(flat["site_quat"][sid], [1, 0, 0, 0])
        b = flat["site_bodyid"][sid]
        assert b == flat["body_name"].index(f"ee{s}_site")
        sb.append(b)
        # every masked joint must lie on the site's chain (so the site Jacobian column is the joint's own)
        anc = set()
        d = lastdof[b]
        while d >= 0:
            anc.add(d)
            d = flat["dof_parentid"][d]
        assert set(t["arm_mask"][a][: t["arm_nmask"][a]]) <= anc
        # pad -> arm ownership
    t["arm_sitebody"] = sb
    t["arm_mocap"] = [K.MOCAP_ID_R, K.MOCAP_ID_L][:narm]
    pad_arm = []
    for pb in t["pad_body"]:
        hand = flat["body_parent"][pb]
        owner = -1
        for a in range(narm):
            x = sb[a]
            while x > 0:
                if x == hand:
                    owner = a
                x = flat["body_parent"][x]
        assert owner >= 0
        pad_arm.append(owner)
    t["pad_arm"] = pad_arm
    return t


def _arr(name: str, vals: List[int], ctype: str = "int") -> str:
    return f"  static constexpr {ctype} {name}[{max(len(vals), 1)}] = {{{', '.join(str(v) for v in (vals or [0]))}}};"


def render_scene_header(scene: str, flat: Dict) -> str:
    """Text of csrc/scenes/scene_<scene>.h for a flat model."""
    t = topology(scene, flat)
    S = SCENE_STRUCT[scene]
    lines = [
        f"// GENERATED by gym_kmanip_b200/scenegen.py from assets/flat/{scene}.json -- do not edit.",
        "// Compile-time topology of one scene; numeric model parameters are run-time kernel arguments.",
        "#pragma once",
        f"struct {S} {{",
        f"  static constexpr int SCENE_ID = {SCENE_ID[scene]};",
    ]
    for k in ["NBODY", "NQ", "NV", "NU", "NVA", "NQA", "NMOCAP", "Q_LEN", "CUBE_BODY", "TABLE_BODY", "NFRIC", "NPAD",
              "TABLE_GEOM", "CUBE_GEOM", "NARM"]:
        lines.append(f"  static constexpr int {k} = {t[k]};")
    for k in ["body_parent", "body_jtype", "body_qadr", "body_dadr", "body_moving", "body_lastdof", "body_chain",
              "body_mocap", "body_hasmass", "dof_parent", "dof_body", "fric_dof", "pad_geom", "pad_body", "pad_arm",
              "arm_nmask", "arm_sitebody", "arm_mocap"]:
        lines.append(_arr(k, list(t[k])))
    lines.append("  static constexpr int arm_mask[%d][8] = {%s};" % (
        t["NARM"], ", ".join("{" + ", ".join(map(str, m)) + "}" for m in t["arm_mask"])))
    lines.append("  static constexpr int arm_grip[%d][2] = {%s};" % (
        t["NARM"], ", ".join("{" + ", ".join(map(str, m)) + "}" for m in t["arm_grip"])))
    lines.append("};")
    return "\n".join(lines) + "\n"


def write_scene_header(scene: str, flat: Dict) -> str:
    os.makedirs(SCENE_DIR, exist_ok=True)
    path = os.path.join(SCENE_DIR, f"scene_{scene}.h")
    with open(path, "w") as f:
        f.write(render_scene_header(scene, flat))
    return path
