"""Shared helpers of the parity tests: the relative-error metric and oracle-side batch rollouts."""
import numpy as np

from oracle import oracle as orc_mod


def rel_err(a, b, floor=1e-3):
    """max |a - b| / max(|b|, floor) over one field of one batch (the scale is the field's own magnitude)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor))


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
