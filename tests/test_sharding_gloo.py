"""CPU, world_size 2 over gloo: the multi-GPU host logic.  Each rank owns a contiguous range of global env ids
(sharding.shard_range), advances it independently (here with the oracle standing in for the device, since no GPU is
attached), and the only communication is the single all-reduce of the rollout totals (sharding.reduce_totals).
Results must not depend on the number of ranks: the cube spawn is keyed by the global env id."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_kmanip_b200 import sharding

TOTAL, STEPS, SEED = 10, 66, 21


def _rollout(env0, n, total_seed=SEED):
    from oracle import oracle as om
    o = om.Oracle("KManipSoloArm")
    st = om.batch_reset_state(o, n, seed=total_seed, env0=env0)
    if o.nmocap == 0:
        st["mocap"] = np.zeros((n, 0))
    totals = np.zeros(4)
    last_obs = None
    for t in range(STEPS):
        # per-env action streams keyed by the global env id, as a sharded policy would produce them
        a = np.stack([np.random.default_rng([SEED, env0 + i, t]).uniform(-1, 1, o.task.act_dim) for i in range(n)]).astype(np.float32)
        obs, fobs, rew, trunc, flags, ncon, geoms = om.batch_step(o, st, a, autoreset=True, seed=total_seed, env0=env0, nthreads=1)
        totals += [rew.sum(), n, trunc.sum(), (rew > 2.0).sum()]
        last_obs = obs
    return totals, last_obs, st["qpos"].copy()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env0, n = sharding.shard_range(TOTAL, rank, world)
    totals, obs, qpos = _rollout(env0, n)
    t = torch.from_numpy(totals.copy())
    dist.barrier()
    sharding.reduce_totals(t)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), totals=t.numpy(), obs=obs, qpos=qpos, env0=env0, n=n)
    dist.destroy_process_group()


def test_two_rank_rollout_equals_single_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref_totals, ref_obs, ref_qpos = _rollout(0, TOTAL)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for p in parts:                     # every rank holds the job-wide totals after the one collective
        assert np.allclose(p["totals"], ref_totals, rtol=1e-12)
    assert ref_totals[1] == TOTAL * STEPS and ref_totals[2] == TOTAL   # one truncation per env in 66 steps
    obs = np.concatenate([p["obs"] for p in parts])
    qpos = np.concatenate([p["qpos"] for p in parts])
    assert np.array_equal(obs, ref_obs) and np.array_equal(qpos, ref_qpos)   # bit-identical, independent of sharding
    assert [int(p["env0"]) for p in parts] == [0, 5]


def test_reduce_totals_is_a_noop_without_a_group():
    t = torch.tensor([1.0, 2.0], dtype=torch.float64)
    assert sharding.reduce_totals(t) is t and t.tolist() == [1.0, 2.0]
