"""Completed KManip model <-> MuJoCo.

The scenes of the reference cannot be compiled as shipped (every robot / table geom is a git-ignored STL mesh and no
robot body has an <inertial>), so this package flattens them with its own loader (mjcf.py) and the
model-completion spec (assets/completion_spec.json).  This module closes the loop with the real engine where it is
installed:

  completed_mjcf(flat)      the completed model as a self-contained MJCF string (primitives, explicit <inertial>s,
                            explicit <contact><pair>s) that ``mujoco.MjModel.from_xml_string`` compiles;
  flat_from_mjmodel(m)      the flat-model dict (the km_model arrays of include/kmanip_b200.h) filled from a compiled
                            ``mujoco.MjModel`` -- the path BASELINE.json's north_star names ("compiles the scene with
                            mujoco.MjModel only to flatten model arrays") and INTEGRATION.md 1 describes.

tests/test_mujoco_conformance.py uses both (auto-skipping without ``mujoco``): the arrays MuJoCo computes -- including
the compile-time constants dof_invweight0 / body_invweight0 / stat.meaninertia that mjcf.py restates -- are compared
with mjcf.py's, and the oracle is stepped against ``mj_step`` on the same model.  Nothing here is on the hot path.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

_JNT = {0: "free", 1: "ball", 2: "slide", 3: "hinge"}
_GEOM = {0: "plane", 2: "sphere", 3: "capsule", 5: "cylinder", 6: "box"}


def _f(v) -> str:
    return " ".join(repr(float(x)) for x in np.atleast_1d(np.asarray(v, dtype=np.float64)))


def completed_mjcf(flat: Dict) -> str:
    """MJCF of the completed model: same body / joint / dof / geom / pair / actuator / site order as `flat`."""
    nb = flat["nbody"]
    children: List[List[int]] = [[] for _ in range(nb)]
    for b in range(1, nb):
        children[flat["body_parent"][b]].append(b)
    sites_of = [[s for s in range(flat["nsite"]) if flat["site_bodyid"][s] == b] for b in range(nb)]
    geoms_of = [[g for g in range(flat["ngeom"]) if flat["geom_bodyid"][g] == b] for b in range(nb)]
    cams_of = [[c for c in range(len(flat.get("cam_name", []))) if flat["cam_bodyid"][c] == b] for b in range(nb)]
    opt = flat["opt"]
    out = ['<mujoco model="kmanip_completed_%s">' % flat.get("scene", "scene").replace(".xml", "")]
    out.append('  <compiler angle="radian" autolimits="false" inertiafromgeom="false" boundmass="0" boundinertia="0"/>')
    out.append('  <option timestep="%s" gravity="%s" tolerance="%s" iterations="%d" ls_iterations="%d" ls_tolerance="%s" impratio="%s" '
               'integrator="Euler" solver="Newton" cone="pyramidal" jacobian="dense"/>'
               % (_f(opt["timestep"]), _f(opt["gravity"]), _f(opt["tolerance"]), opt["iterations"], opt["ls_iterations"],
                  _f(opt["ls_tolerance"]), _f(opt["impratio"])))

    def emit_body_contents(b: int, ind: str):
        if b > 0 and flat["body_mass"][b] > 0:
            out.append('%s<inertial pos="%s" mass="%s" diaginertia="%s"/>' % (ind, _f(flat["body_ipos"][b]), _f(flat["body_mass"][b]),
                                                                               _f(flat["body_inertia"][b])))
        for k in range(flat["body_jntnum"][b]):
            j = flat["body_jntadr"][b] + k
            jt = flat["jnt_type"][j]
            if jt == 0:
                d = flat["jnt_dofadr"][j]
                out.append('%s<joint name="%s" type="free" frictionloss="%s" solreffriction="%s" solimpfriction="%s"/>'
                           % (ind, flat["jnt_name"][j], _f(flat["dof_frictionloss"][d]), _f(flat["dof_solref"][d]), _f(flat["dof_solimp"][d])))
                continue
            d = flat["jnt_dofadr"][j]
            out.append('%s<joint name="%s" type="%s" pos="%s" axis="%s" limited="%s" range="%s" solreflimit="%s" solimplimit="%s" '
                       'frictionloss="%s" solreffriction="%s" solimpfriction="%s" armature="0" damping="0" stiffness="0"/>'
                       % (ind, flat["jnt_name"][j], _JNT[jt], _f(flat["jnt_pos"][j]), _f(flat["jnt_axis"][j]),
                          "true" if flat["jnt_limited"][j] else "false", _f(flat["jnt_range"][j]), _f(flat["jnt_solref"][j]),
                          _f(flat["jnt_solimp"][j]), _f(flat["dof_frictionloss"][d]), _f(flat["dof_solref"][d]), _f(flat["dof_solimp"][d])))
        for g in geoms_of[b]:
            gt = flat["geom_type"][g]
            size = flat["geom_size"][g]
            size = size[:1] if gt == 2 else (size if gt in (0, 6) else size[:2])
            out.append('%s<geom name="%s" type="%s" size="%s" pos="%s" quat="%s" contype="0" conaffinity="0" mass="0"/>'
                       % (ind, flat["geom_name"][g], _GEOM[gt], _f(size), _f(flat["geom_pos"][g]), _f(flat["geom_quat"][g])))
        for s in sites_of[b]:
            out.append('%s<site name="%s" pos="%s" quat="%s"/>' % (ind, flat["site_name"][s], _f(flat["site_pos"][s]), _f(flat["site_quat"][s])))
        for c in cams_of[b]:
            tgt = flat["cam_targetbodyid"][c]
            mode = ' mode="targetbody" target="%s"' % flat["body_name"][tgt] if tgt >= 0 else ""
            out.append('%s<camera name="%s" pos="%s" fovy="%s"%s/>' % (ind, flat["cam_name"][c], _f(flat["cam_pos"][c]), _f(flat["cam_fovy"][c]), mode))
        for c in children[b]:
            mocap = ' mocap="true"' if flat["body_mocapid"][c] >= 0 else ""
            out.append('%s<body name="%s" pos="%s" quat="%s"%s>' % (ind, flat["body_name"][c], _f(flat["body_pos"][c]), _f(flat["body_quat"][c]), mocap))
            emit_body_contents(c, ind + "  ")
            out.append("%s</body>" % ind)

    out.append("  <worldbody>")
    emit_body_contents(0, "    ")
    out.append("  </worldbody>")
    out.append("  <contact>")
    for p in range(flat["npair"]):
        out.append('    <pair geom1="%s" geom2="%s" condim="%d" friction="%s" solref="%s" solimp="%s" margin="%s" gap="%s"/>'
                   % (flat["geom_name"][flat["pair_geom1"][p]], flat["geom_name"][flat["pair_geom2"][p]], flat["pair_condim"][p],
                      _f(flat["pair_friction"][p]), _f(flat["pair_solref"][p]), _f(flat["pair_solimp"][p]), _f(flat["pair_margin"][p]),
                      _f(flat["pair_gap"][p])))
    out.append("  </contact>")
    out.append("  <actuator>")
    for a in range(flat["nu"]):
        out.append('    <position name="act%d" joint="%s" kp="%s" ctrllimited="%s" ctrlrange="%s" forcelimited="%s" forcerange="%s"/>'
                   % (a, flat["jnt_name"][flat["act_jntid"][a]], _f(flat["act_kp"][a]), "true" if flat["act_ctrllimited"][a] else "false",
                      _f(flat["act_ctrlrange"][a]), "true" if flat["act_forcelimited"][a] else "false", _f(flat["act_forcerange"][a])))
    out.append("  </actuator>")
    out.append("</mujoco>")
    return "\n".join(out)


# keys of the flat model that are plain numeric arrays and must agree between mjcf.py and a compiled MjModel
NUMERIC_KEYS = ["body_parent", "body_rootid", "body_mocapid", "body_pos", "body_quat", "body_mass", "body_ipos", "body_inertia",
                "body_invweight0", "body_jntadr", "body_jntnum", "jnt_type", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_pos",
                "jnt_axis", "jnt_limited", "jnt_range", "jnt_solref", "jnt_solimp", "qpos0", "dof_bodyid", "dof_jntid", "dof_parentid",
                "dof_frictionloss", "dof_solref", "dof_solimp", "dof_invweight0", "act_jntid", "act_kp", "act_ctrllimited",
                "act_ctrlrange", "act_forcelimited", "act_forcerange", "site_bodyid", "site_pos", "site_quat", "geom_type",
                "geom_bodyid", "geom_pos", "geom_quat", "geom_size", "pair_geom1", "pair_geom2", "pair_condim", "pair_friction",
                "pair_solref", "pair_solimp", "pair_margin", "pair_gap", "mocap_pos0", "mocap_quat0"]


def flat_from_mjmodel(m, template: Dict = None) -> Dict:
    """The flat-model dict of this package filled from a compiled ``mujoco.MjModel`` (names follow mjModel).  `template`
    supplies the entries MuJoCo does not carry for this path (scene name, render appearance)."""
    import mujoco
    nm = lambda kind, i: mujoco.mj_id2name(m, kind, i) or ""   # noqa: E731
    L = lambda a: np.asarray(a).tolist()                        # noqa: E731
    nmocap = int(m.nmocap)
    mocap_bodies = [b for b in range(m.nbody) if m.body_mocapid[b] >= 0]
    mocap_bodies.sort(key=lambda b: m.body_mocapid[b])
    flat = dict(template or {})
    flat.update(
        nbody=int(m.nbody), njnt=int(m.njnt), nq=int(m.nq), nv=int(m.nv), nu=int(m.nu), nsite=int(m.nsite), ngeom=int(m.ngeom),
        npair=int(m.npair), nmocap=nmocap,
        opt=dict(timestep=float(m.opt.timestep), gravity=L(m.opt.gravity), tolerance=float(m.opt.tolerance), iterations=int(m.opt.iterations),
                 ls_iterations=int(m.opt.ls_iterations), ls_tolerance=float(m.opt.ls_tolerance), impratio=float(m.opt.impratio)),
        meaninertia=float(m.stat.meaninertia),
        body_name=[nm(mujoco.mjtObj.mjOBJ_BODY, b) for b in range(m.nbody)],
        body_parent=L(m.body_parentid), body_rootid=L(m.body_rootid), body_mocapid=L(m.body_mocapid), body_pos=L(m.body_pos),
        body_quat=L(m.body_quat), body_mass=L(m.body_mass), body_ipos=L(m.body_ipos), body_inertia=L(m.body_inertia),
        body_invweight0=L(m.body_invweight0), body_jntadr=L(m.body_jntadr), body_jntnum=L(m.body_jntnum),
        jnt_name=[nm(mujoco.mjtObj.mjOBJ_JOINT, j) for j in range(m.njnt)],
        jnt_type=L(m.jnt_type), jnt_bodyid=L(m.jnt_bodyid), jnt_qposadr=L(m.jnt_qposadr), jnt_dofadr=L(m.jnt_dofadr), jnt_pos=L(m.jnt_pos),
        jnt_axis=L(m.jnt_axis), jnt_limited=L(m.jnt_limited.astype(int)), jnt_range=L(m.jnt_range), jnt_solref=L(m.jnt_solref),
        jnt_solimp=L(m.jnt_solimp), qpos0=L(m.qpos0),
        dof_bodyid=L(m.dof_bodyid), dof_jntid=L(m.dof_jntid), dof_parentid=L(m.dof_parentid), dof_frictionloss=L(m.dof_frictionloss),
        dof_solref=L(m.dof_solref), dof_solimp=L(m.dof_solimp), dof_invweight0=L(m.dof_invweight0),
        act_jntid=L(m.actuator_trnid[:, 0]), act_kp=L(m.actuator_gainprm[:, 0]), act_ctrllimited=L(m.actuator_ctrllimited.astype(int)),
        act_ctrlrange=L(m.actuator_ctrlrange), act_forcelimited=L(m.actuator_forcelimited.astype(int)), act_forcerange=L(m.actuator_forcerange),
        site_name=[nm(mujoco.mjtObj.mjOBJ_SITE, s) for s in range(m.nsite)],
        site_bodyid=L(m.site_bodyid), site_pos=L(m.site_pos), site_quat=L(m.site_quat),
        geom_name=[nm(mujoco.mjtObj.mjOBJ_GEOM, g) for g in range(m.ngeom)],
        geom_type=L(m.geom_type), geom_bodyid=L(m.geom_bodyid), geom_pos=L(m.geom_pos), geom_quat=L(m.geom_quat), geom_size=L(m.geom_size),
        pair_geom1=L(m.pair_geom1), pair_geom2=L(m.pair_geom2), pair_condim=L(m.pair_dim), pair_friction=L(m.pair_friction),
        pair_solref=L(m.pair_solref), pair_solimp=L(m.pair_solimp), pair_margin=L(m.pair_margin), pair_gap=L(m.pair_gap),
        mocap_pos0=[L(m.body_pos[b]) for b in mocap_bodies], mocap_quat0=[L(m.body_quat[b]) for b in mocap_bodies],
    )
    return flat
