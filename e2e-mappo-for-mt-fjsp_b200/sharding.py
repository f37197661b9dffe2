"""Env-slice sharding across the GPUs of one box.

Environments are independent, so the batch is cut into contiguous slices: env i of the job lives on rank
i // envs_per_rank.  There is no data-path collective.  The only exchanges are the ones the north star names for the
rollout side: the max-over-ranks step time (benchmarks) and the episode statistics (sum of the 4 final costs, the
weighted objective and the env count: 6 doubles, one allreduce per episode).  Works with any torch.distributed
backend (nccl on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import instances


def shard_range(total_envs: int, rank: int, world: int):
    """Contiguous slice [first, first+count) of rank `rank`; the first `total % world` ranks get one more env."""
    base, rem = divmod(total_envs, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def owner_of(env: int, total_envs: int, world: int) -> int:
    base, rem = divmod(total_envs, world)
    cut = rem * (base + 1)
    return env // (base + 1) if env < cut else rem + (env - cut) // max(base, 1)


def make_shard(total_envs, rank, world, n_job, n_machine, n_edge, seed, episode=0):
    """Instance slice + reward weights of this rank, identical to the same rows of the unsharded batch."""
    first, count = shard_range(total_envs, rank, world)
    d = instances.synthetic_instances(first, count, n_job, n_machine, n_edge, seed)
    w = instances.random_weights(first, count, seed, episode)
    return first, count, d, w


def _dev(t, device):
    return t.to(device) if device is not None else t


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = _dev(torch.tensor([value], dtype=torch.float64), device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_episode_stats(cost4, weights=(0.4, 0.4, 0.2), device=None):
    """Mean final costs over ALL envs of the job: cost4 [b,4] = (mk, pt/N, tt, idle) of this rank's slice
    (Run.py:632-640 averages over the env batch; validate.py:283 defines the objective).
    Returns dict(mk, pt, tt, idle, objective, count)."""
    c = torch.as_tensor(cost4, dtype=torch.float64)
    obj = weights[0] * c[:, 0] + weights[1] * (c[:, 1] + c[:, 3]) + weights[2] * c[:, 2]
    s = torch.cat([c.sum(0), obj.sum(0, keepdim=True), torch.tensor([float(c.shape[0])], dtype=torch.float64,
                                                                    device=c.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        s = _dev(s, device)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    s = s.cpu().numpy()
    n = s[5]
    return dict(mk=s[0] / n, pt=s[1] / n, tt=s[2] / n, idle=s[3] / n, objective=s[4] / n, count=int(n))
