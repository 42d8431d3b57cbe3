"""Multi-GPU plumbing: envs are independent, so ranks own contiguous env ranges and the step has
NO collective.  ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used
only for the rollout statistics -- the counterpart of what rsl_rl's logger averages from
``extras["log"]`` (reference: scripts/rsl_rl/train.py:125-132 for the per-rank device / seed)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import native


def shard_env_range(global_num_envs: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the envs rank ``rank`` owns (contiguous, sizes differ by at most one)."""
    base, rem = divmod(int(global_num_envs), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_seed(seed: int, rank: int) -> int:
    """``seed + local_rank`` (scripts/rsl_rl/train.py:130-132)."""
    return int(seed) + int(rank)


def reduce_rollout_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce one 32-float statistics slot across ranks.

    Words 0..15 are per-rank MEANS over the envs reset on that rank (``Episode_Reward/<term>``):
    they are turned into sums with the reset count (word 16), summed over ranks and divided again, so
    the result equals what a single process holding all envs would log.  Counters (16..21) add up."""
    s = stats.detach().clone().to(torch.float64)
    n = s[native.STAT_NUM_RESET].clone()
    s[:native.MAX_TERMS] *= n
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    tot = s[native.STAT_NUM_RESET]
    s[:native.MAX_TERMS] = torch.where(tot > 0, s[:native.MAX_TERMS] / tot.clamp(min=1), torch.zeros_like(s[:native.MAX_TERMS]))
    return s.to(stats.dtype)
