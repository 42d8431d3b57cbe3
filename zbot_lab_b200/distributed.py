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


class RolloutStatsReducer:
    """The same reduction, OFF the stepping stream: ``submit(stats)`` snapshots the slot on the caller's stream and runs the
    all-reduce on a side stream, so the next control steps overlap it (the statistics of a rollout are only read by the
    logger after the rollout); ``result()`` joins the side stream back and returns the last reduced slot.  On CPU tensors
    (gloo tests) it degrades to the synchronous call."""

    def __init__(self, device, group=None):
        self.device = torch.device(device)
        self.group = group
        self.side = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._last = None
        self.submitted = 0

    def submit(self, stats: torch.Tensor) -> None:
        self.submitted += 1
        if self.side is None:
            self._last = reduce_rollout_stats(stats, self.group)
            return
        cur = torch.cuda.current_stream(self.device)
        snap = stats.detach().clone()                 # on the stepping stream: the ring slot may be rewritten 64 steps later
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            snap.record_stream(self.side)
            self._last = reduce_rollout_stats(snap, self.group)

    def result(self):
        if self.side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        return self._last
