#!/usr/bin/env python
"""Regenerates the committed golden fixtures of tests/golden/.

    python tests/golden/make_golden.py

What pins what (the reference's own tests hold no numeric vectors):
  scipy_rotation.json   scipy.spatial.transform.Rotation on seeded inputs: the Euler / quaternion conventions the
                        reference uses at env_sim.py:62-66 (as_euler("xyz"), from_euler("xyz").as_quat()[[3,0,1,2]])
                        and the rotation-vector difference of ik_mujoco.py:43-46 (mju_subQuat).
  ik_trf.json           the reference's ik() (ik_mujoco.py:100-155) run with the genuine scipy.optimize.least_squares
                        on the oracle's restated ik_res / ik_jac, for seeded goals; pins the batched device IK.
  fk_home.json          end-effector site positions at the home pose computed during the survey by an independent
                        throw-away script -- pins the MJCF flattening conventions.
  traj_<env>.npz        teacher-forcing records of the oracle (state before, action, outputs, state after): pins the
                        oracle against accidental change and is what the CUDA path is compared with on the GPU box
                        without needing to run the oracle there.
  render_<env>.npz      camera observations of the Vision ids: four mid-episode states of the trajectory fixture, the
                        oracle's render records (oracle/render_oracle.py: camera frame + primitive list from the C++
                        oracle's body frames) and its ray-cast images (head camera at 160 x 120, gripper cameras at their
                        own 60 x 40); pins the numpy restatement and is what km_render is compared with on the GPU box.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as om   # noqa: E402
from parity_util import oracle_rollout, pack_state   # noqa: E402

TRAJ_ENVS = ["KManipSoloArm", "KManipSoloArmQPos", "KManipDualArm", "KManipDualArmQPos", "KManipTorso"]
TRAJ_N, TRAJ_STEPS = 16, 66


def scipy_rotation():
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(7)
    rots = R.random(24, random_state=11)
    mats = rots.as_matrix()
    eul = rots.as_euler("xyz")
    quat_wxyz = R.from_euler("xyz", eul).as_quat()[:, [3, 0, 1, 2]]
    d_eul = eul + rng.uniform(-0.1, 0.1, eul.shape)
    quat_goal = R.from_euler("xyz", d_eul).as_quat()[:, [3, 0, 1, 2]]
    # mju_subQuat(qa, qb): rotation vector of qb^-1 * qa expressed in qb's frame
    ra, rb = R.random(24, random_state=5), R.random(24, random_state=6)
    sub = (rb.inv() * ra).as_rotvec()
    return dict(mat=mats.tolist(), euler_xyz=eul.tolist(), quat_wxyz=quat_wxyz.tolist(), euler_goal=d_eul.tolist(),
                quat_goal_wxyz=quat_goal.tolist(), qa_wxyz=ra.as_quat()[:, [3, 0, 1, 2]].tolist(),
                qb_wxyz=rb.as_quat()[:, [3, 0, 1, 2]].tolist(), subquat=sub.tolist())


def ik_trf():
    out = {}
    for env_id in ("KManipSoloArm", "KManipDualArm", "KManipTorso"):
        o = om.Oracle(env_id)
        rng = np.random.default_rng(3)
        cases = []
        st0 = om.batch_reset_state(o, 1, seed=0)
        for c in range(6):
            qpos = st0["qpos"][0].copy()
            for a in range(o.task.n_arm):
                for i in range(o.task.arm_nmask[a]):
                    j = o.task.arm_mask[a][i]
                    lo, hi = o.flat["jnt_range"][j]
                    qpos[j] = np.clip(qpos[j] + rng.uniform(-0.2, 0.2), lo + 1e-3, hi - 1e-3)
            o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
            for a in range(o.task.n_arm):
                sid = o.task.arm_site[a]
                spos = o.field("site_xpos").reshape(-1, 3)[sid]
                smat = o.field("site_xmat").reshape(-1, 9)[sid]
                from scipy.spatial.transform import Rotation as R
                eul = R.from_matrix(smat.reshape(3, 3)).as_euler("xyz") + rng.uniform(-1, 1, 3).astype(np.float32) * 0.1
                gq = R.from_euler("xyz", eul).as_quat()[[3, 0, 1, 2]]
                gp = spos + rng.uniform(-1, 1, 3).astype(np.float32) * 0.01
                o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
                q_trf = o.ik_trf(a, gp, gq, qpos)
                o.set_state(qpos, np.zeros(o.nv), qpos[: o.nu])
                q_dls = o.ik_dls(a, gp, gq, qpos)
                cases.append(dict(arm=a, qpos=qpos.tolist(), goal_pos=gp.tolist(), goal_quat=gq.tolist(),
                                  q_trf=q_trf.tolist(), q_dls=q_dls.tolist()))
        out[env_id] = cases
    return out


def fk_home():
    # SURVEY.md section 8c, "[survey-derived, throw-away script, intrinsic-xyz euler, not MuJoCo]"
    return {
        "KManipSoloArm": {"eer_site_pos": {"pos": [0.25766, 0.49943, 0.62639], "quat": [0.92933, -0.10107, -0.08387, -0.34511]}},
        "KManipDualArm": {"eer_site_pos": {"pos": [0.25792, 0.49897, 0.62654]}, "eel_site_pos": {"pos": [-0.17239, 0.57768, 0.65785]}},
        "KManipTorso": {"eer_site_pos": {"pos": [0.18345, 0.41856, 0.55541]}, "eel_site_pos": {"pos": [-0.17732, 0.40926, 0.52584]}},
    }


def trajectories():
    for env_id in TRAJ_ENVS:
        o, traj = oracle_rollout(env_id, TRAJ_N, TRAJ_STEPS, seed=11, action_seed=13)
        keep = [0, 1, 2, 20, 40, 62, 63, 64, 65]   # early free flight, mid episode, the truncation / autoreset boundary
        rec = {}
        for t in keep:
            r = traj[t]
            rec[f"s{t}_before"] = pack_state(r["before"])
            rec[f"s{t}_after"] = pack_state(r["after"])
            for k in ("step", "episode"):
                rec[f"s{t}_before_{k}"] = r["before"][k]
                rec[f"s{t}_after_{k}"] = r["after"][k]
            for k in ("action", "obs", "final_obs", "reward", "truncated", "flags", "ncon", "geoms"):
                rec[f"s{t}_{k}"] = r[k]
        rec["steps"] = np.array(keep)
        np.savez_compressed(os.path.join(HERE, f"traj_{env_id}.npz"), **rec)


RENDER_ENVS = {"KManipSoloArm": ["head", "grip_r"], "KManipDualArm": ["head", "grip_l", "grip_r"], "KManipTorso": ["head", "grip_l", "grip_r"]}
RENDER_SIZE = {"head": (160, 120), "grip_r": (60, 40), "grip_l": (60, 40)}


def render_fixtures():
    from oracle import render_oracle as ro
    for env_id, cams in RENDER_ENVS.items():
        traj = np.load(os.path.join(HERE, f"traj_{env_id}.npz"))
        states = traj["s40_after"][:4]
        o = om.Oracle(env_id)
        nq, nv, nu = o.nq, o.nv, o.nu
        rec = {"state": states}
        frames = []
        for st in states:
            o.set_state(st[:nq], st[nq:nq + nv], st[nq + nv:nq + nv + nu])
            o.mj_forward(disable_actuation=True)
            frames.append((o.field("xpos").reshape(-1, 3), o.field("xquat").reshape(-1, 4)))
        for cam in cams:
            w, h = RENDER_SIZE[cam]
            P = ro.params(o.flat, cam, w, h)
            recs = np.stack([ro.scene_record(o.flat, xp, xq, cam) for xp, xq in frames])
            rec[f"rec_{cam}"] = recs
            rec[f"img_{cam}"] = np.stack([ro.render_record(r, P) for r in recs])
        np.savez_compressed(os.path.join(HERE, f"render_{env_id}.npz"), **rec)


def main():
    om.build()
    if "--render-only" in sys.argv:
        render_fixtures()
        return
    json.dump(scipy_rotation(), open(os.path.join(HERE, "scipy_rotation.json"), "w"))
    json.dump(ik_trf(), open(os.path.join(HERE, "ik_trf.json"), "w"))
    json.dump(fk_home(), open(os.path.join(HERE, "fk_home.json"), "w"), indent=1)
    trajectories()
    render_fixtures()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
