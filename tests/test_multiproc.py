"""World-size-2 CPU test (gloo) of the only cross-rank logic on the path: env sharding, per-rank
seeds and the rollout-statistics reduction (there is no collective inside the step)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zbot_lab_b200 import distributed as zd
    lo, hi = zd.shard_env_range(1001, rank, world)
    # per-rank statistics slot: means over this rank's reset envs + counters
    rng = np.random.default_rng(100 + rank)
    nreset = 3 + 4 * rank
    sums = rng.normal(size=(nreset, 13))
    slot = torch.zeros(32)
    slot[:13] = torch.from_numpy(sums.mean(0)).float()
    slot[16] = nreset
    slot[17] = 1 + rank
    slot[18] = nreset - 1 - rank
    slot[19] = float(rank + 0.5)
    out = zd.reduce_rollout_stats(slot)
    q.put((rank, lo, hi, sums, out.numpy(), zd.rank_seed(42, rank)))
    dist.destroy_process_group()


def test_stats_reduction_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, o0, seed0), (r1, lo1, hi1, s1, o1, seed1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)          # contiguous, complete, balanced
    assert (seed0, seed1) == (42, 43)
    assert np.array_equal(o0, o1)                               # every rank ends with the same slot
    want = np.concatenate([s0, s1]).mean(0)                      # what one process with all envs would log
    assert np.allclose(o0[:13], want, atol=1e-6)
    assert o0[16] == 3 + 7 and o0[17] == 1 + 2 and o0[18] == (3 - 1) + (7 - 2) and o0[19] == 2.0


def test_shard_ranges_cover_everything():
    from zbot_lab_b200.distributed import shard_env_range
    for n in (1, 7, 4096, 65536 * 8 + 3):
        for w in (1, 2, 4, 8):
            r = [shard_env_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
