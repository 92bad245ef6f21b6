"""Shared helpers of the parity tests: the relative-error metric and oracle-side batch rollouts."""
import numpy as np

from oracle import oracle as orc_mod


def rel_err(a, b, floor=1e-3):
    """max |a - b| / max(|b|, floor) over one field of one batch (the scale is the field's own magnitude)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor))


# Per-component relative error: every component is compared on the scale of its own unit (a batch-max over a
# record that mixes radians, metres of slider travel and a unit quaternion hides a 1e-3 relative error on a slider).
# floor = the magnitude below which a component's error is measured absolutely (its typical scale).
UNIT_FLOOR = dict(hinge=0.1, slide=0.01, cube_pos=0.1, quat=1.0,          # rad, m (0.03 m travel), m, -
                  hinge_vel=1.0, slide_vel=0.01, cube_lin=0.1, cube_ang=1.0)  # rad/s, m/s, m/s, rad/s


def component_floors(flat, nq, nv):
    """Per-component floors for a qpos [nq] and a qvel [nv] record of a KManip scene (links first, free cube last)."""
    jt = list(flat["jnt_type"])[: nv - 6]      # 3 hinge, 2 slide
    fq = np.array([UNIT_FLOOR["hinge"] if t == 3 else UNIT_FLOOR["slide"] for t in jt] + [UNIT_FLOOR["cube_pos"]] * 3 + [UNIT_FLOOR["quat"]] * 4)
    fv = np.array([UNIT_FLOOR["hinge_vel"] if t == 3 else UNIT_FLOOR["slide_vel"] for t in jt] + [UNIT_FLOOR["cube_lin"]] * 3 + [UNIT_FLOOR["cube_ang"]] * 3)
    assert fq.size == nq and fv.size == nv
    return fq, fv


def comp_rel_err(a, b, floors):
    """max over envs and components of |a - b| / max(|b|, floor of the component); a, b: [n, k]."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floors[None, :])))


def comp_rel_err_per_env(a, b, floors):
    """[n] per-env maximum over components of |a - b| / max(|b|, floor of the component)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floors[None, :]), axis=1)


def rel_err_per_env(a, b, floor=1e-3):
    """[n] per-env max |a - b| on the scale of the field's batch magnitude (the per-env resolution of rel_err)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b), axis=1) / max(float(np.max(np.abs(b))), floor)


def pack_state(st):
    """Batch state dict -> [n, state_dim] records (include/kmanip_b200.h layout; cube_lo = 0)."""
    n = st["qpos"].shape[0]
    return np.concatenate([st["qpos"], st["qvel"], st["ctrl"], st["warm"], st["mocap"][:, : max(st["mocap"].shape[1], 0)],
                           st["time"].reshape(n, 1), np.zeros((n, 3))], axis=1)


# fp32 build, envs whose cube is in contact: absolute bounds on one env step.  Typical errors are 2e-5 m / 2e-3 s^-1;
# the bound is set by the rare cube balancing on ONE corner, an unstable equilibrium that amplifies rounding
# differences within the 10 sub-steps (DESIGN.md "fp32 and the cube").
CONTACT_TOL_POS_F32 = 1e-3
CONTACT_TOL_VEL_F32 = 0.2
_REAL_FIELDS = ("qpos", "qvel", "ctrl", "warm", "mocap", "time")


def oracle_rollout(env_id, n, steps, seed=0, action_seed=1, nthreads=0, round32=False, **okw):
    """Free-running oracle rollout of n envs; returns the list of (state_before, action, outputs, state_after).
    round32: round the state to float32-representable values before every step, so that an fp32 build can be handed
    exactly the oracle's input (teacher forcing without an input-rounding error)."""
    o = orc_mod.Oracle(env_id, **okw)
    st = orc_mod.batch_reset_state(o, n, seed=seed)
    if o.nmocap == 0:
        st["mocap"] = np.zeros((n, 0))
    rng = np.random.default_rng(action_seed)
    out = []
    for _ in range(steps):
        a = rng.uniform(-1, 1, (n, o.task.act_dim)).astype(np.float32)
        if round32:
            for k in _REAL_FIELDS:
                st[k][...] = st[k].astype(np.float32).astype(np.float64)
        before = {k: v.copy() for k, v in st.items()}
        peak = np.zeros(n, dtype=np.int32)
        obs, fobs, rew, trunc, flags, ncon, geoms = orc_mod.batch_step(o, st, a, autoreset=True, seed=seed, nthreads=nthreads,
                                                                       ncon_peak=peak)
        after = {k: v.copy() for k, v in st.items()}
        out.append(dict(before=before, action=a, obs=obs, final_obs=fobs, reward=rew, truncated=trunc, flags=flags, ncon=ncon,
                        geoms=geoms, after=after, ncon_peak=peak))
    return o, out
