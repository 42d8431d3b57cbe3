"""Distributed PPO iteration on N ranks of one box: where the time goes, and what the gradient all-reduce costs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29534 \
        tools/bench_ppo_ranks.py [--envs 4096] [--iters 10] [--out gpurun_out/ppo_ranks_N.json]

The reference's multi-GPU training (scripts/rsl_rl/train.py:125-132: one process per GPU, ``seed + local_rank``) leaves the
collective to rsl_rl: parameters broadcast once, gradients all-reduced after every mini-batch backward.  ``OnPolicyRunner``
here does the same (one flat all-reduce per mini-batch + the adaptive-schedule KL scalar); this script times it with CUDA
events around every ``dist.all_reduce`` of the update, next to the rollout (one CUDA-graph replay of 24 fused steps) and
the whole learn phase.  Rank 0 prints one JSON object (max over ranks for the phase times).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    from zbot_lab_b200.rl import ppo_runner
    from zbot_lab_b200 import distributed as zdist

    cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "env_cfg_entry_point")
    cfg.scene.num_envs, cfg.sim.device, cfg.seed = args.envs, str(dev), zdist.rank_seed(1, rank)
    env = RslRlVecEnvWrapper(gym.make("zbot-6b-walking-v2", cfg=cfg, render_mode=None))
    acfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "rsl_rl_cfg_entry_point").to_dict()
    acfg["use_cuda_graph"] = True
    runner = ppo_runner.OnPolicyRunner(env, acfg, log_dir=None, device=str(dev))

    # time every all-reduce of the update with CUDA events (the runner calls dist.all_reduce through its module's `dist`)
    events, nbytes = [], [0]
    real_all_reduce = dist.all_reduce

    class _Dist:
        def __getattr__(self, k):
            return getattr(dist, k)

        @staticmethod
        def all_reduce(t, *a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = real_all_reduce(t, *a, **kw)
            e1.record()
            events.append((e0, e1))
            nbytes[0] += t.numel() * t.element_size()
            return r

    if world > 1:
        ppo_runner.dist = _Dist()

    runner.learn(3, init_at_random_ep_len=True)            # warm-up: graph capture, NCCL communicators, allocator
    events.clear()
    nbytes[0] = 0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    hist = runner.learn(args.iters)[-args.iters:]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ar_ms = sum(a.elapsed_time(b) for a, b in events)
    t = torch.tensor([wall, sum(h["collection_s"] for h in hist), sum(h["learn_s"] for h in hist), ar_ms * 1e-3],
                     device=dev, dtype=torch.float64)
    if world > 1:
        real_all_reduce(t, op=dist.ReduceOp.MAX)
    wall, coll, learn, ar = (float(x) for x in t)
    if rank == 0:
        n_param = sum(p.numel() for p in runner.policy.parameters())
        out = {
            "n_ranks": world, "envs_per_rank": args.envs, "iterations": args.iters, "rollout_steps": runner.num_steps,
            "policy_parameters": n_param,
            "ms_per_iteration": 1e3 * wall / args.iters, "collection_ms_per_iteration": 1e3 * coll / args.iters,
            "learn_ms_per_iteration": 1e3 * learn / args.iters,
            "allreduce_ms_per_iteration": 1e3 * ar / args.iters, "allreduce_calls_per_iteration": len(events) / args.iters,
            "allreduce_bytes_per_iteration": nbytes[0] / args.iters,
            "allreduce_share_of_iteration": (ar / wall) if wall > 0 else None,
            "env_steps_per_s_whole_job": world * args.envs * runner.num_steps * args.iters / wall,
            "note": "phase times: max over ranks; all-reduce: CUDA events around every dist.all_reduce of the update "
                    "(flat gradient per mini-batch + the KL scalar of the adaptive schedule)",
        }
        s = json.dumps(out)
        print(s, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            with open(args.out, "w") as f:
                f.write(s + "\n")
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
