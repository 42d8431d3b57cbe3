"""Concurrent PCIe ceiling and end-to-end step on 1..N ranks of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/e2e_ranks.py [--bind] [--out gpurun_out/e2e_ranks_N.json]

Every rank owns one GPU and measures, SIMULTANEOUSLY with the others (barrier before every section):

* pinned D2H / H2D copy bandwidth at the byte counts of one 65536-env step (100 B/env out, 24 B/env in) -- CUDA events;
* D2H and H2D at the same time on two streams (the step's two directions overlap inside the fused launch);
* ``env.step_host`` (the e2e path of bench.py: zero-copy actions in / rows out inside the one kernel launch) -- wall clock
  around 200 synchronous steps, max over ranks;

and reports where it sits: the GPU's PCI address and NUMA node (sysfs), the CPUs the process may run on, and -- with
``--bind`` -- pins the process to the CPUs of the GPU's NUMA node BEFORE any pinned allocation (first touch then places the
pinned pages on that node).  Rank 0 prints one JSON object: per-rank numbers, the aggregate, and the e2e step as a fraction of
the concurrent D2H ceiling (rows_bytes / d2h_GBps = the floor of a step whose results must cross PCIe).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_pci_numa(index: int):
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        short = bus.lower()
        if len(short.split(":")[0]) == 8:
            short = short[4:]
        node = None
        p = f"/sys/bus/pci/devices/{short}/numa_node"
        if os.path.exists(p):
            node = int(open(p).read().strip())
        try:
            gen = pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h)
            width = pynvml.nvmlDeviceGetCurrPcieLinkWidth(h)
        except Exception:
            gen = width = None
        return short, node, gen, width
    except Exception as exc:   # noqa: BLE001
        return repr(exc), None, None, None


def node_cpus(node: int):
    p = f"/sys/devices/system/node/node{node}/cpulist"
    if not os.path.exists(p):
        return None
    cpus = []
    for part in open(p).read().strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    return cpus


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true", help="pin the process to the CPUs of the GPU's NUMA node before allocating")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bus, node, gen, width = gpu_pci_numa(local)
    bound = None
    if args.bind and node is not None and node >= 0:
        cpus = node_cpus(node)
        allowed = sorted(set(cpus or []) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            bound = len(allowed)

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.envs
    out_b, in_b = n * 100, n * 24
    res = {"rank": rank, "pci": bus, "numa_node": node, "pcie_gen": gen, "pcie_width": width,
           "cpus_allowed": len(os.sched_getaffinity(0)), "bound_to_node_cpus": bound}

    h_out = torch.empty(out_b, dtype=torch.uint8).pin_memory()
    d_out = torch.empty(out_b, dtype=torch.uint8, device=dev)
    h_in = torch.empty(in_b, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(in_b, dtype=torch.uint8, device=dev)

    def timed_copies(pairs, reps=50):
        """pairs: [(dst, src, stream)], all issued per repetition; returns us per repetition (events on the default stream
        bracketing, the side streams joined back)."""
        s0 = torch.cuda.current_stream()
        for _ in range(5):
            for dst, src, s in pairs:
                with torch.cuda.stream(s):
                    dst.copy_(src, non_blocking=True)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s0)
        for dst, src, s in pairs:
            s.wait_stream(s0)
        for _ in range(reps):
            for dst, src, s in pairs:
                with torch.cuda.stream(s):
                    dst.copy_(src, non_blocking=True)
        for dst, src, s in pairs:
            s0.wait_stream(s)
        b.record(s0)
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / reps

    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    us = timed_copies([(h_out, d_out, s1)])
    res["d2h_rows_us"], res["d2h_GBps"] = us, out_b / us / 1e3
    us = timed_copies([(d_in, h_in, s1)])
    res["h2d_actions_us"], res["h2d_GBps"] = us, in_b / us / 1e3
    us = timed_copies([(h_out, d_out, s1), (d_in, h_in, s2)])
    res["both_directions_us"] = us
    big = 64 << 20
    hb = torch.empty(big, dtype=torch.uint8).pin_memory()
    db = torch.empty(big, dtype=torch.uint8, device=dev)
    us = timed_copies([(hb, db, s1)], reps=10)
    res["d2h_64MiB_GBps"] = big / us / 1e3
    del hb, db

    # the e2e path of bench.py, all ranks at once
    from zbot_lab_b200.compat import gym_registry as gym
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.utils import synthetic as syn
    import numpy as np

    cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "env_cfg_entry_point")
    cfg.scene.num_envs = n
    cfg.sim.device = str(dev)
    cfg.seed = 1234 + rank
    env = gym.make("zbot-6b-walking-v2", cfg=cfg, render_mode=None)
    env.reset()
    st = env.unwrapped._stepper
    rng = np.random.default_rng(1234 + rank)
    st.set_sim_state({k: torch.from_numpy(v).to(dev) for k, v in syn.synth_sim_state(rng, n).items()})
    h_act = [torch.randn(n, 6).pin_memory() for _ in range(4)]
    h_rows = torch.empty(n, 25).pin_memory()
    d_act = torch.randn(n, 6, device=dev)
    for i in range(20):
        env.step_host(h_act[i % 4], h_rows)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        env.step_host(h_act[i % 4], h_rows)
    torch.cuda.synchronize()
    res["step_host_us"] = (time.perf_counter() - t0) / args.steps * 1e6
    barrier()
    # the same kernel, device-resident, synchronised every step (what the host path adds is the difference)
    for i in range(20):
        env.step(d_act)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        env.step(d_act)
        torch.cuda.synchronize()
    res["step_device_sync_us"] = (time.perf_counter() - t0) / args.steps * 1e6
    barrier()
    env.close()

    allres = [None] * world
    if world > 1:
        dist.all_gather_object(allres, res)
    else:
        allres = [res]
    if rank == 0:
        worst = max(r["step_host_us"] for r in allres)
        agg_d2h = sum(r["d2h_GBps"] for r in allres)
        floor_us = max(out_b / (r["d2h_GBps"] * 1e3) for r in allres)
        summary = {
            "n_ranks": world, "envs_per_rank": n, "bind": bool(args.bind), "rows_bytes_per_step": out_b, "action_bytes_per_step": in_b,
            "aggregate_d2h_GBps_concurrent": agg_d2h, "min_rank_d2h_GBps": min(r["d2h_GBps"] for r in allres),
            "aggregate_h2d_GBps_concurrent": sum(r["h2d_GBps"] for r in allres),
            "step_host_us_max_over_ranks": worst, "e2e_env_steps_per_s": world * n / (worst * 1e-6),
            "d2h_floor_us_slowest_rank": floor_us, "e2e_fraction_of_d2h_ceiling": floor_us / worst,
            "step_device_sync_us_max_over_ranks": max(r["step_device_sync_us"] for r in allres),
            "ranks": allres,
        }
        s = json.dumps(summary)
        print(s, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            with open(args.out, "w") as f:
                f.write(s + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
