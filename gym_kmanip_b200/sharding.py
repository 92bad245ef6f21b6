"""Env sharding across GPUs / ranks: contiguous global env-id ranges, no per-step communication.

The hot path is embarrassingly parallel over envs: each rank owns the envs
[env0, env0 + n) of the job, runs the identical kernels on its own GPU, and the cube-spawn generator is keyed
by the *global* env id, so results do not depend on the number of ranks.  The only collective of a rollout is
one all-reduce (sum) of a handful of float64 totals at its end (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(env0, n) of `rank`: contiguous ranges, the first `total % world` ranks take one extra env."""
    if not (0 <= rank < world_size) or total_envs < 0:
        raise ValueError("bad rank / world_size / total_envs")
    base, extra = divmod(total_envs, world_size)
    n = base + (1 if rank < extra else 0)
    env0 = rank * base + min(rank, extra)
    return env0, n


def owner_of(global_env: int, total_envs: int, world_size: int) -> Tuple[int, int]:
    """(rank, local index) that owns a global env id under shard_range."""
    base, extra = divmod(total_envs, world_size)
    split = extra * (base + 1)
    if global_env < split:
        return global_env // (base + 1), global_env % (base + 1)
    if base == 0:
        raise ValueError("global_env out of range")
    r = extra + (global_env - split) // base
    return r, (global_env - split) % base


def reduce_totals(totals, group=None):
    """Sum a small float64 tensor of rollout totals over all ranks (in place); a no-op without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)
    return totals
