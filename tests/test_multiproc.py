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
    red = zd.RolloutStatsReducer("cpu")          # the overlapped form degrades to the synchronous call on CPU tensors
    red.submit(slot)
    assert torch.equal(red.result(), out) and red.submitted == 1
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


def _ppo_worker(rank, world, port, q):
    """Two ranks, each with its own envs (different seeds -> different rollouts) and its own copy of the policy: after
    `learn` every rank must hold the SAME parameters (rank-0 broadcast at construction + flat-gradient all-reduce per
    mini-batch = what rsl_rl does under --distributed; reference scripts/rsl_rl/train.py:125-132)."""
    import sys
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import zbot_lab_b200.tasks  # noqa: F401
    import zbot_lab_b200.tasks.zbot6b_direct.walking_v2 as w2
    from fake_stepper import FakeStepper
    from zbot_lab_b200 import distributed as zd
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner
    w2.NativeStepper = FakeStepper                       # CPU double of the stepper (this box has no GPU)
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "env_cfg_entry_point")
    cfg.scene.num_envs, cfg.sim.device, cfg.seed = 8, "cpu", zd.rank_seed(11, rank)
    env = RslRlVecEnvWrapper(gym.make("zbot-6b-walking-v2", cfg=cfg, render_mode=None))
    torch.manual_seed(1000 + rank)                       # different initial weights per rank: the broadcast must fix that
    acfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "rsl_rl_cfg_entry_point").to_dict()
    acfg["num_steps_per_env"] = 4
    r = OnPolicyRunner(env, acfg, log_dir=None, device="cpu")
    assert r.distributed
    flat0 = torch.cat([p.detach().reshape(-1) for p in r.policy.parameters()]).clone()
    r.learn(2)
    flat1 = torch.cat([p.detach().reshape(-1) for p in r.policy.parameters()])
    rew = float(r.buf["rew"].sum())
    q.put((rank, flat0.numpy(), flat1.numpy(), rew, r.lr))
    dist.destroy_process_group()


def test_distributed_ppo_ranks_end_with_identical_parameters_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ppo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, a0, a1, rew0, lr0), (_, b0, b1, rew1, lr1) = res
    assert np.array_equal(a0, b0)                 # broadcast from rank 0 at construction
    assert np.array_equal(a1, b1)                 # identical after two PPO iterations (all-reduced gradients, shared KL -> shared lr)
    assert not np.array_equal(a0, a1) and rew0 != rew1 and lr0 == lr1      # they did learn, on different rollouts
