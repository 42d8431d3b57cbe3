#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused ``zbot-6b-walking-v2`` step on N B200s of one node.

  python bench.py                         # N=1: 65536 envs (headline `value`) + 4096 envs ("envs_4096",
                                          # BASELINE.json configs[1]) + the MDP-only kernel side measurement
  torchrun ... bench.py --gpus 8 ...      # 65536 envs/GPU, env-sharded (configs[2]); NCCL only
                                          # for the rollout statistics
  python bench.py --impl reference        # the CPU path (oracle CPU port, all host threads)

One "step" = one control step of every env (4 physics substeps + MDP + partial reset) = ONE
kernel launch.  Prints one JSON line (rank 0).  Timing: CUDA events on the launching stream
around every block of 50 control steps, inputs larger than L2 (the steps cycle through independent env sets that together
exceed 2 x the L2; `--l2 flush` = the round-1 method, a 256 MiB write fill before every step), max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec zbot-6b-walking-v2 fused step"
UNIT = "env-steps/s"
# ALGORITHMIC HBM bytes per env-step of the fused step: SURVEY.md §8(d) (read 336 B + write 394 B of SoA f32 state,
# actions, observation, reward, flags) -- the figure `roofline.achieved` / `roofline.frac` are computed with.  The bytes
# the kernel's actual LAYOUT moves (80-word padded state read + written, int64 episode counter) are
# reported next to it as `layout_bytes_per_env_step` / `frac_layout_bytes`.
ALGO_BYTES_PER_ENV_STEP = 730
LAYOUT_BYTES_PER_ENV_STEP = 320 + 24 + 8 + 320 + 92 + 4 + 2 + 8
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 148 SMs x 128 FP32 lanes x 2 FLOP (FMA) x 1.965 GHz = 74.4
# Envs per GPU of the headline `value` at EVERY N (weak scaling: identical per-GPU work at N = 1, 2, 4, 8 so the
# driver's scaling efficiency is meaningful).  65536 envs/GPU is the configuration BASELINE.json states the
# multi-GPU target on (configs[2]); the 4096-env configuration (configs[1]) is measured in the same run at N = 1
# and reported under "envs_4096".
DEFAULT_ENVS_PER_GPU = 65536
ROLLOUT_STEPS = 24  # agents/rsl_rl_ppo_cfg.py:67 -- statistics are reduced once per rollout
# One bench "step" = a BLOCK of this many control steps (each one launch of the fused kernel on the next env set of the
# rotation: inputs larger than L2), so the driver's `--steps 20` times 1000 control steps (>= 50 ms of
# kernel time) instead of 1.7 ms.  `value` counts every control step: env-steps/s is unaffected by the block size.
CONTROL_STEPS_PER_STEP = 50
L2_BYTES = 126 << 20


def workload_name(n_envs: int) -> str:
    """The ONE workload string both arms (`--impl ours` / `--impl reference`) print in `config.workload`."""
    return f"zbot-6b-walking-v2 full control step (4 physics substeps + MDP + partial reset), {n_envs} envs/GPU"


def csrc_hash() -> str:
    """Content hash of the sources the library is built from: a committed ncu summary (profiles/step_<N>.json) is only
    used when it was captured on exactly these sources."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "zbot_lab_b200", "csrc", "*.cu")) +
                    glob.glob(os.path.join(ROOT, "zbot_lab_b200", "csrc", "*.h")) +
                    glob.glob(os.path.join(ROOT, "zbot_lab_b200", "csrc", "*.cuh"))) + [os.path.join(ROOT, "include", "zbot_b200.h")]:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def load_step_profile(n_envs: int):
    """profiles/step_<N>.json (written by tools/ncu_profile_json.py from one `ncu --set full --clock-control none` capture
    of the step kernel at N envs): instruction counts, DRAM bytes, duration.  Returns (dict | None, note)."""
    p = os.path.join(ROOT, "profiles", f"step_{n_envs}.json")
    if not os.path.isfile(p):
        return None, f"no committed ncu summary for {n_envs} envs (profiles/step_{n_envs}.json)"
    try:
        prof = json.load(open(p))
    except Exception as exc:
        return None, f"unreadable {p}: {exc!r}"
    if prof.get("source_hash") != csrc_hash():
        return None, (f"profiles/step_{n_envs}.json was captured on other kernel sources (hash {prof.get('source_hash')} != "
                      f"{csrc_hash()}): stale, not used")
    return prof, "profiles/step_%d.json (%s)" % (n_envs, prof.get("captured_with", "ncu"))


def l2_note(m) -> str:
    if m["l2_mode"] == "rotate":
        return ("inputs larger than L2: %d independent env sets of %d envs (%.0f MB of state, counters and outputs in all = %.1f x "
                "the 126 MB L2; 16 rotating action buffers) stepped in turn, back to back, no flush; one CUDA-event pair per "
                "block of %d control steps" % (m["n_sets"], m["n_envs"], m["n_sets"] * m["set_bytes"] / 1e6,
                                                m["n_sets"] * m["set_bytes"] / L2_BYTES, CONTROL_STEPS_PER_STEP))
    if m["l2_mode"] == "flush":
        return "flushed before every control step (256 MiB write, outside the timed events, events around every control step)"
    return "not flushed: one env set back to back (state stays in L2)"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=20)
    p.add_argument("--envs", type=int, default=None, help="envs per GPU (default 65536 at every N)")
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--l2", default="rotate", choices=["rotate", "flush", "none"],
                   help="rotate: cycle through env sets that together exceed 2 x L2 (default); flush: 256 MiB write fill before "
                        "every control step; none: one env set back to back (state stays in L2)")
    p.add_argument("--no-flush", action="store_true", help="alias of --l2 none")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-mdp", action="store_true", help="skip the MDP-only kernel side measurement")
    p.add_argument("--no-small", action="store_true", help="skip the additional 4096-env measurement at N=1")
    p.add_argument("--no-tasks", action="store_true", help="skip the snake / v4 / PPO-rollout side measurements at N=1")
    p.add_argument("--cpu-sample-steps", type=int, default=None)
    a = p.parse_args()
    if a.no_flush:
        a.l2 = "none"
    return a


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._stop_evt = threading.Event()

    def sample_once(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                  "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
            f = [x.strip() for x in out.split(",")]
            self.samples.append({"sm": float(f[0]), "max": float(f[1]), "power": float(f[2]),
                                 "hw_slowdown": f[3], "hw_thermal": f[4], "sw_thermal": f[5], "sw_power_cap": f[6]})
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample_once()
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(s["sm"] for s in self.samples)
        reasons = sorted({k for s in self.samples for k in ("hw_slowdown", "hw_thermal", "sw_thermal", "sw_power_cap")
                          if s[k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0]["max"], "reasons": reasons,
                "samples": len(self.samples), "power_w_max": max(s["power"] for s in self.samples)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_path(n_envs: int, steps: int, warmup: int, seed: int = 1234):
    """The path on the host cores: oracle/cpu_port (float32, OpenMP over envs)."""
    import numpy as np

    from oracle import cpu_port
    from zbot_lab_b200.utils import synthetic as syn

    cores = cpu_port.set_threads(None)          # all host threads, even under torchrun (OMP_NUM_THREADS=1)
    rng = np.random.default_rng(seed)
    env = cpu_port.PortEnv(n_envs, np.float32)
    env.set_sim_state(syn.synth_sim_state(rng, n_envs))
    env.ep_len[:] = rng.integers(0, 1000, n_envs)
    acts = rng.normal(0, 1, (8, n_envs, 6)).astype(np.float32)
    for i in range(warmup):
        env.step(acts[i % 8])
    t0 = time.perf_counter()
    for i in range(steps):
        env.step(acts[i % 8])
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; this arm must use every host thread, and libgomp
    # only behaves (thread pool, spinning) when it sees the final settings at process start: re-exec once.
    want = str(os.cpu_count() or 1)
    if os.environ.get("OMP_NUM_THREADS") != want and os.environ.get("ZBOT_BENCH_REEXEC") != "1":
        env = dict(os.environ, OMP_NUM_THREADS=want, ZBOT_BENCH_REEXEC="1")
        sys.stdout.flush()
        os.execve(sys.executable, [sys.executable] + sys.argv, env)
    n_envs = args.envs or DEFAULT_ENVS_PER_GPU
    # Same metric / config / steps / warm-up as the GPU arm.  One bench step is a block of control steps; the block is
    # the "bounded sample": CONTROL_STEPS_PER_STEP control steps when the whole run then fits ~2 minutes of CPU time,
    # fewer otherwise (env-steps/s is a rate: the block length does not enter `value`).
    steps, warm = max(1, args.steps), max(0, args.warmup)
    _, dt_probe, cores = cpu_path(n_envs, 2, 1)
    per_ctrl = dt_probe / 2
    block = int(max(1, min(CONTROL_STEPS_PER_STEP, 120.0 / (per_ctrl * (steps + warm)))))
    value, dt, cores = cpu_path(n_envs, steps * block, warm * block)
    sample = (f"{n_envs} envs x {steps} bench steps x {block} control steps (+ {warm} x {block} warm-up), "
              f"oracle/cpu_port.cpp float32, OpenMP, {dt:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(n_envs), "envs_per_gpu": n_envs},
        "timing": {"control_steps_per_step": block, "runs_on": f"host CPU, {cores} threads", "wall_s": dt},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "kind = port: the reference's step is torch eager MDP + Isaac Lab + PhysX (closed, not installable here or on "
                "the GPU box), so this arm times the C++ CPU port of the SAME control step (oracle/cpu_port.cpp = "
                "csrc/zbot_core.h compiled for the host, OpenMP over envs, all host threads) -- NOT the reference's own "
                "torch / PhysX CPU path, and a faster baseline than it (the reference's torch MDP alone, without physics, "
                "ran at 1.3e6 env-steps/s on 8 vCPU: BASELINE.md §1; see cpu_baseline_torch_mdp in the GPU arm's line when "
                "a reference tree is present)",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist

    from zbot_lab_b200.compat import gym_registry as gym
    import zbot_lab_b200.tasks  # noqa: F401  (registers zbot-6b-walking-v2)
    from zbot_lab_b200 import distributed as zdist
    from zbot_lab_b200.utils import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the zbot step has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_envs = args.envs or DEFAULT_ENVS_PER_GPU

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(n_envs, steps, warmup, with_e2e, mode=None):
        """Time `steps` control steps of `n_envs` envs on this rank; returns a dict (rank-local + max-reduced)."""

        # the public API: gym.make(task, cfg=...) exactly as scripts/rsl_rl/train.py:158
        def make_set(k):
            cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "env_cfg_entry_point")
            cfg.scene.num_envs = n_envs
            cfg.sim.device = str(dev)
            cfg.seed = zdist.rank_seed(1234 + 7919 * k, rank)
            env_k = gym.make("zbot-6b-walking-v2", cfg=cfg, render_mode=None)
            st_k = env_k.unwrapped._stepper
            env_k.reset()
            rng = np.random.default_rng(1234 + rank + 7919 * k)
            st_k.set_sim_state({kk: torch.from_numpy(v).to(dev) for kk, v in syn.synth_sim_state(rng, n_envs).items()})
            g_k = torch.Generator(device=dev).manual_seed(1234 + rank + 7919 * k)
            env_k.episode_length_buf = torch.randint(0, 1000, (n_envs,), device=dev, generator=g_k)
            return env_k, st_k

        # L2 policy (bench contract: flush between timed iterations OR inputs larger than L2).  "rotate" (default): the control
        # steps cycle through `n_sets` independent env sets whose state + outputs together are >= 2 x the 126 MB L2, so every
        # step finds its state in DRAM and the steps still run back to back on the stream (one CUDA-event pair per block of
        # 50 control steps).  "flush": ONE env set, a 256 MiB write fill before every control step, events around each step
        # (the round-1 method; the fill leaves the L2 full of dirty lines that the step then has to write back).
        mode = mode or args.l2
        set_bytes = n_envs * (LAYOUT_BYTES_PER_ENV_STEP - 320)       # state + counters + actions + outputs of one env set
        n_sets = min(96, max(2, -(-2 * L2_BYTES // set_bytes))) if mode == "rotate" else 1
        sets = [make_set(k) for k in range(n_sets)]
        env, st = sets[0]
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        n_act = 16
        actions = torch.randn(n_act, n_envs, 6, device=dev, generator=g)            # resident in HBM
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if mode == "flush" else None
        stats_acc = torch.zeros(32, device=dev)
        reducer = zdist.RolloutStatsReducer(dev)

        def one_step(i):
            sets[i % n_sets][1].step(actions[i % n_act])

        def launch_count():
            return sum(s_k.launch_count for _, s_k in sets)

        blk = CONTROL_STEPS_PER_STEP
        ctrl_steps, ctrl_warm = steps * blk, warmup * blk
        for i in range(ctrl_warm):
            one_step(i)
            if world > 1 and (i + 1) % ROLLOUT_STEPS == 0:
                reducer.submit(sets[i % n_sets][1].stats)            # warm-up covers the collective too (NCCL channel set-up)
        barrier()
        sampler = ClockSampler(local)
        sampler.sample_once()
        sampler.start()
        per = 1 if flush is not None else blk                        # control steps per CUDA-event pair
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ctrl_steps // per)]
        # the only collective on the path: the rollout statistics, all-reduced once per rollout (24 control steps) -- inside
        # the timed region, with its own events (N > 1 only)
        n_red = ctrl_steps // ROLLOUT_STEPS if world > 1 else 0
        ev_red = [torch.cuda.Event(enable_timing=True) for _ in range(n_red)]
        ev_sub = [torch.cuda.Event(enable_timing=True) for _ in range(n_red)]
        launches0 = launch_count()
        barrier()
        t_wall0 = time.perf_counter()
        k_red = 0
        for i in range(ctrl_steps):
            if flush is not None:
                flush.fill_(i & 0xFF)                      # evict L2 (126 MB) -- outside the timed events
            if i % per == 0:
                ev[i // per][0].record()
            one_step(i)
            if k_red < n_red and (i + 1) % ROLLOUT_STEPS == 0:
                # on a side stream, overlapped with the next control steps (the logger reads it after the rollout); its own
                # events are recorded on that stream, and whatever of the LAST one outlives the last step is added below
                ev_sub[k_red].record()
                reducer.submit(sets[i % n_sets][1].stats)
                with torch.cuda.stream(reducer.side):
                    ev_red[k_red].record()
                k_red += 1
            if (i + 1) % per == 0:
                ev[i // per][1].record()                    # after the submit: its 128-byte snapshot copy is inside the timed events
        ev_tail = torch.cuda.Event(enable_timing=True)
        if n_red:
            stats_acc = reducer.result()                   # joins the side stream into the stepping stream
        ev_tail.record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = launch_count() - launches0
        sampler.stop()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        # all-reduce k is submitted right after control step 24(k+1)-1: its latency = that point of the stepping stream -> its end
        red_ms = [ev_sub[k].elapsed_time(ev_red[k]) for k in range(n_red)]
        tail_ms = max(0.0, ev[-1][1].elapsed_time(ev_tail)) if n_red else 0.0
        total_ms = float(sum(step_ms)) + tail_ms
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        value = world * n_envs * ctrl_steps / (total_ms * 1e-3)
        steps_e2e = ctrl_steps

        # end to end through the public API with HOST buffers: pinned actions -> H2D, env.step, D2H of the result
        e2e = None
        if with_e2e:
            h_act = [torch.randn(n_envs, 6).pin_memory() for _ in range(4)]
            h_out = torch.empty(n_envs * 98, dtype=torch.uint8).pin_memory()   # obs | rew | terminated | truncated
            h_rows = torch.empty(n_envs, 25).pin_memory()                      # step_host rows: obs | rew | flags
            d_act = torch.empty(n_envs, 6, device=dev)

            def e2e_step_staged(i):
                d_act.copy_(h_act[i % 4], non_blocking=True)          # this step's inputs: pinned host -> device
                env.step(d_act)                                       # the public API call
                h_out.copy_(env.last_step_packed, non_blocking=True)  # the step's whole result: device -> pinned host
                torch.cuda.synchronize()
                return env.unpack_host(h_out)

            def e2e_step(i):
                # the host-facing public API call: pinned actions in, packed result out, both zero-copy over PCIe
                # inside the one kernel launch; returns after the stream is synchronised (result owned by the host)
                return env.step_host(h_act[i % 4], h_rows)

            for i in range(max(3, ctrl_warm // 4)):
                e2e_step(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(steps_e2e):
                e2e_step(i)
            barrier()
            te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e = {"value": world * n_envs * steps_e2e / float(te.item()), "unit": UNIT,
                   "h2d_bytes_per_step": n_envs * 24, "d2h_bytes_per_step": n_envs * 100,
                   "path": "env.step_host -> zbot_step_host: pinned host actions in, (N,25) pinned host rows "
                           "(obs | reward | flags) out, both zero-copy over PCIe inside the one fused-kernel launch; "
                           "synchronous (stream synchronised every step)"}
            obs_h, rew_h, term_h, trunc_h = e2e_step(0)
            assert obs_h.shape == (n_envs, 23) and bool(torch.isfinite(rew_h).all())
            # the staged variant (explicit H2D copy, device-resident step, one packed D2H copy) for comparison
            for i in range(3):
                e2e_step_staged(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(steps_e2e):
                e2e_step_staged(i)
            barrier()
            ts = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            e2e["staged_copies_value"] = world * n_envs * steps_e2e / float(ts.item())
            # context only (NOT the headline): a device-resident consumer, as in the reference's own RL loop where the policy
            # lives on the GPU -- pinned host actions -> device every step, env.step, and a device -> host read of ONE scalar
            # (the step's reward sum from the statistics slot); shows how much of `value` -> `e2e` is the 100 B/env of results
            # crossing PCIe
            def e2e_metric_only(i):
                d_act.copy_(h_act[i % 4], non_blocking=True)
                env.step(d_act)
                return float(st.stats[19].item())
            for i in range(3):
                e2e_metric_only(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(steps_e2e):
                e2e_metric_only(i)
            barrier()
            tm = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            e2e["metric_only_value"] = world * n_envs * steps_e2e / float(tm.item())
            e2e["metric_only_note"] = ("context: H2D of the actions + env.step + D2H of one scalar per step (device-resident "
                                       "observations, as with a GPU policy); h2d %d B, d2h 4 B per step" % (n_envs * 24))


        kernel_name = st.kernel_name
        for env_k, _ in sets:
            env_k.close()
        return {"n_envs": n_envs, "value": value, "total_ms": total_ms, "ctrl_steps": ctrl_steps, "kernel": kernel_name,
                "local_ms_per_ctrl_step": float(sum(step_ms)) / ctrl_steps,
                "stats_allreduce": ({"count": len(red_ms), "avg_us": 1e3 * float(sum(red_ms)) / len(red_ms),
                                     "median_us": 1e3 * float(sorted(red_ms)[len(red_ms) // 2]), "max_us": 1e3 * float(max(red_ms)),
                                     "every_control_steps": ROLLOUT_STEPS, "words": int(stats_acc.numel()),
                                     "overlapped": "side stream, concurrent with the following control steps; latency = end of the "
                                                   "submitting step -> end of the all-reduce",
                                     "tail_ms_added_to_timed_region": tail_ms,
                                     "included_in_value": True} if red_ms else None),
                "launches": int(launches), "e2e": e2e, "clocks": sampler.summary(), "wall": t_wall, "l2_mode": mode, "n_sets": n_sets,
                "set_bytes": set_bytes}

    main_m = measure(n_envs, args.steps, args.warmup, not args.no_e2e)
    step_kernel_name = main_m["kernel"]
    small_m = None
    if world == 1 and args.envs is None and not args.no_small:
        small_m = measure(4096, args.steps, args.warmup, not args.no_e2e)     # BASELINE.json configs[1]
    value, total_ms, launches, e2e = main_m["value"], main_m["total_ms"], main_m["launches"], main_m["e2e"]
    flush_m = None
    if world == 1 and args.l2 == "rotate" and not args.no_small:
        # the round-1 / early round-2 method beside it, for continuity: one env set, 256 MiB write fill before every control step
        flush_m = measure(n_envs, max(2, min(4, args.steps)), 1, False, mode="flush")

    mdp_only = None
    if rank == 0 and world == 1 and not args.no_mdp:
        # the HBM-bound first kernel (BASELINE.json configs[0] shape): MDP-only step on synthetic articulation state
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_mdp

        peak_m, _ = measured_peak_gbs()
        mdp_only = {}
        for n_m in (65536, 262144):
            # inputs larger than L2 (six rotating input sets, 534 MB at 65536 envs), 12 steps per CUDA-graph replay, no flush
            ms = bench_mdp.bench_mdp(n_m, 20, mode="graph-rotate")
            gbs = bench_mdp.MDP_ALGO_BYTES * n_m / (ms * 1e-3) / 1e9
            # the round-1 method beside it: a 256 MB write fill between steps leaves the L2 full of dirty lines whose write-back
            # (126 MB) shares the DRAM with the step's 108 MB of reads -- it bounds the fraction at ~0.44 for ANY kernel at 65536
            ms_w = bench_mdp.bench_mdp(n_m, 30, 5, mode="write")
            gbs_w = bench_mdp.MDP_ALGO_BYTES * n_m / (ms_w * 1e-3) / 1e9
            mdp_only[str(n_m)] = {"ms_per_step": ms, "env_steps_per_s": n_m / (ms * 1e-3), "achieved_gbs": gbs,
                                  "frac_of_measured_hbm_peak": gbs / peak_m,
                                  "l2": "inputs larger than L2 (6 rotating input sets), CUDA-graph replay of 12 steps, no flush",
                                  "after_256MB_write_flush": {"ms_per_step": ms_w, "achieved_gbs": gbs_w,
                                                              "frac_of_measured_hbm_peak": gbs_w / peak_m}}
        mdp_only["kernel"] = ("zbot_mdp_kernel<true>, 112-env tiles (65536 envs = two full waves of 2 x 148 CTAs), statistics fused "
                              "into its last CTA: one launch per step")
        mdp_only["algorithmic_bytes_per_env_step"] = bench_mdp.MDP_ALGO_BYTES

    other_tasks = None
    if rank == 0 and world == 1 and args.envs is None and not args.no_tasks:
        # the other BASELINE.json configs, device-resident, back to back after warm-up (CUDA events around the loop)
        from zbot_lab_b200 import native
        from zbot_lab_b200.stepper import NativeStepper

        def time_task(task, n_t, steps_t=100):
            if task == native.TASK_WALKING_M:
                terms = [(f, w, p) for _, f, w, p in native.M_FLAT_TERMS if f != "is_terminated"]
                st_t = NativeStepper(n_t, dev, native.make_m_cfg(n_t, terms, is_terminated_weight=-200.0, act_clip=0.04 * 3.141592653589793))
            else:
                st_t = NativeStepper(n_t, dev, native.make_cfg(n_t, task=task))
            if task == native.TASK_WALKING_M:
                st_t.reset_idx_m(None)
                st_t.state.set("joint_speed_limit", torch.rand(n_t, 1, device=dev) * 0.7 + 0.3)      # friction 0.3 .. 1.0
            elif task == native.TASK_WALKING_V4:
                st_t.reset_idx_v4(None)
                st_t.state.set("base_pos_y_err_sum", torch.rand(n_t, 1, device=dev) * 3.0 + 3.0)
            else:
                st_t.reset_idx(None)
            gt = torch.Generator(device=dev).manual_seed(7)
            st_t.episode_length_buf[:] = torch.randint(0, 790, (n_t,), device=dev, generator=gt)
            if task == native.TASK_SNAKE_V0:
                st_t.state.set("joint_speed_limit", (torch.rand(n_t, 1, device=dev, generator=gt) * 1.8 + 0.2) * 3.14159265)
            acts_t = torch.randn(8, n_t, 6, device=dev, generator=gt)
            for i in range(20):
                st_t.step(acts_t[i % 8])
            torch.cuda.synchronize()
            a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = st_t.launch_count
            a_ev.record()
            for i in range(steps_t):
                st_t.step(acts_t[i % 8])
            b_ev.record()
            torch.cuda.synchronize()
            ms = a_ev.elapsed_time(b_ev) / steps_t
            out_t = {"envs": n_t, "ms_per_step": ms, "value": n_t / (ms * 1e-3), "unit": UNIT,
                     "gpu_launches": int(st_t.launch_count - l0), "l2": "not flushed (back to back)"}
            st_t.close()
            return out_t

        def time_no_termination(n_t, steps_t=100):
            """the same fused step with zero actions from the default stance -- nobody falls, nobody resets
            (episode counters start at 0): separates the cost of the reset path from the steady-state step."""
            st_t = NativeStepper(n_t, dev, native.make_cfg(n_t, task=native.TASK_WALKING_V2))
            st_t.reset_idx(None)
            acts_t = torch.zeros(n_t, 6, device=dev)
            for i in range(20):
                st_t.step(acts_t)
            torch.cuda.synchronize()
            a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_ev.record()
            resets = 0.0
            for i in range(steps_t):
                st_t.step(acts_t)
            b_ev.record()
            torch.cuda.synchronize()
            resets = float(st_t.stats_ring[:, native.STAT_NUM_RESET].sum())
            ms = a_ev.elapsed_time(b_ev) / steps_t
            st_t.close()
            return {"envs": n_t, "ms_per_step": ms, "value": n_t / (ms * 1e-3), "unit": UNIT, "resets_in_last_64_steps": resets,
                    "l2": "not flushed (back to back)",
                    "workload": "zbot-6b-walking-v2 fused step, zero actions from the default stance (no terminations, no resets)"}

        other_tasks = {
            "walking_v2_65536_no_termination": time_no_termination(65536),
            "walking_v2_65536_random_actions": dict(time_task(native.TASK_WALKING_V2, 65536),
                                                    workload="zbot-6b-walking-v2 fused step, random actions (falls and resets "
                                                             "every step), back to back"),
            "snake_16384": dict(time_task(native.TASK_SNAKE_V0, 16384),
                                workload="BASELINE.json configs[3]: zbot-6s-snake-v0 fused step, 16384 envs, "
                                         "12 ground spheres in contact per env"),
            "walking_v4_4096": dict(time_task(native.TASK_WALKING_V4, 4096),
                                    workload="zbot-6b-walking-v4 fused step (commands + event resampling, in-kernel RNG), 4096 envs"),
            "walking_v4_65536": dict(time_task(native.TASK_WALKING_V4, 65536),
                                     workload="zbot-6b-walking-v4 fused step, 65536 envs"),
            "walking_m_4096": dict(time_task(native.TASK_WALKING_M, 4096),
                                   workload="zbot-6b-walking-m-v0 (manager-based task, Zbot6BFlatEnvCfg terms, full-inertia robot, "
                                            "per-env friction, in-kernel RNG) fused step, 4096 envs"),
            "walking_m_65536": dict(time_task(native.TASK_WALKING_M, 65536),
                                    workload="zbot-6b-walking-m-v0 fused step, 65536 envs"),
        }
        try:   # BASELINE.json configs[4]: PPO rollout 24 steps x 4096 envs, policy MLP + env step + storage, one CUDA graph
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_rollout
            r = bench_rollout.make(4096, str(dev), True)
            sec = bench_rollout.time_rollouts(r, 30)
            other_tasks["ppo_rollout_24x4096"] = {
                "workload": "BASELINE.json configs[4]: rollout of 24 steps x 4096 envs (actor + critic MLP 3x128 ELU + Gaussian "
                            "sample + log-prob as ONE launch of zbot_policy_act_tc5_kernel (tcgen05.mma kind::tf32, accumulators in TMEM, 3 x TF32 split products = FP32 accuracy), fused env step, zbot_rollout_store) replayed "
                            "as one CUDA graph",
                "ms_per_rollout": 1e3 * sec, "value": 24 * 4096 / sec, "unit": UNIT,
                "policy": "zbot_policy_act (this library)" if r._fused is not None else "torch"}
            r.env.close()
            r = bench_rollout.make(4096, str(dev), True, fused=False)      # the same rollout with the torch policy / storage ops
            sec_t = bench_rollout.time_rollouts(r, 10)
            other_tasks["ppo_rollout_24x4096"]["torch_policy_same_graph_ms"] = 1e3 * sec_t
            r.env.close()
        except Exception as exc:   # the side measurement must never break the headline line
            other_tasks["ppo_rollout_24x4096"] = {"error": repr(exc)}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        kern_ms = main_m["local_ms_per_ctrl_step"]          # one control step = one launch of the step kernel (+ statistics kernel)
        achieved = ALGO_BYTES_PER_ENV_STEP * n_envs / (kern_ms * 1e-3) / 1e9
        achieved_layout = LAYOUT_BYTES_PER_ENV_STEP * n_envs / (kern_ms * 1e-3) / 1e9
        prof, prof_note = load_step_profile(n_envs)
        clk_hz = 1e6 * float((main_m["clocks"] or {}).get("sm_mhz") or 1965.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n_envs), "envs_per_gpu": n_envs},
            "timing": {"control_steps_per_step": CONTROL_STEPS_PER_STEP, "ms_per_control_step": total_ms / main_m["ctrl_steps"],
                       "timed_control_steps": main_m["ctrl_steps"], "timed_region_ms": total_ms,
                       "parallelism": f"env-sharded x{world}, no collective in the step",
                       "l2": l2_note(main_m),
                       "wall_s_incl_flush": main_m["wall"], "stats_allreduce": main_m["stats_allreduce"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (prof["dram_bytes_read"] + prof["dram_bytes_write"]) if prof else None,
                         "peak_source": peak_src, "kernel": step_kernel_name,
                         "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP,
                         "layout_bytes_per_env_step": LAYOUT_BYTES_PER_ENV_STEP, "achieved_layout_bytes": achieved_layout,
                         "frac_layout_bytes": achieved_layout / peak, "profile": prof_note,
                         "note": "algorithmic bytes = SURVEY.md §8(d) x envs per launch / CUDA-event time of the launch; traffic = "
                                 "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (null when the committed "
                                 "summary is stale).  The fused step is FP32-issue bound, not HBM bound (DESIGN.md §4): see fp32_issue"},
            "gpu_launches": int(launches), "clocks": main_m["clocks"],
        }
        if prof:
            # the bound that actually applies: warp instructions issued per clock per SM sub-partition, and the FP32 FLOP rate
            # (2 x FFMA + FMUL + FADD thread instructions of the committed capture / live kernel time) against the FP32 peak
            inst = float(prof["smsp_inst_executed"])
            # scalar FP32 from the sass_thread_inst counters; the packed forms (FFMA2 / FMUL2 / FADD2: two FP32 operations per
            # lane, the packed-halves kernel) only appear in the per-opcode counts of the source page
            packed = [float(prof.get(k, 0.0)) for k in ("op_ffma2", "op_fmul2", "op_fadd2")]
            flop = 32.0 * (2.0 * prof["ffma"] + prof["fmul"] + prof["fadd"] + 2.0 * (2.0 * packed[0] + packed[1] + packed[2]))
            line["fp32_issue"] = {
                "achieved_inst_per_clk_per_smsp": inst / (148 * 4) / (kern_ms * 1e-3 * clk_hz), "nominal": 1.0,
                "warp_instructions_per_launch": inst,
                "fp32_share_of_instructions": (prof["ffma"] + prof["fmul"] + prof["fadd"] + sum(packed)) / inst,
                "packed_fp32_share_of_instructions": sum(packed) / inst,
                "fp32_tflops": flop / (kern_ms * 1e-3) / 1e12, "fp32_peak_tflops": FP32_PEAK_TFLOPS,
                "fp32_peak_frac": flop / (kern_ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS, "sm_clock_mhz": clk_hz / 1e6,
                "registers_per_thread": prof.get("registers"), "warps_per_scheduler": prof.get("warps_per_scheduler"),
                "source": prof_note + ": instruction counts of the committed capture / live CUDA-event kernel time"}
        else:
            line["fp32_issue"] = {"unavailable": prof_note}
        if flush_m is not None:
            line["timing"]["after_256MB_write_flush"] = {
                "value": flush_m["value"], "unit": UNIT, "ms_per_control_step": flush_m["total_ms"] / flush_m["ctrl_steps"],
                "timed_control_steps": flush_m["ctrl_steps"], "l2": l2_note(flush_m),
                "note": "the fill leaves the L2 full of dirty lines; their write-back (126 MB) shares the DRAM with the step, and the "
                        "fill between two steps removes the launch overlap (PDL) the back-to-back stream has"}
        if small_m is not None:
            a4 = ALGO_BYTES_PER_ENV_STEP * 4096 / (small_m["local_ms_per_ctrl_step"] * 1e-3) / 1e9
            line["envs_4096"] = {"workload": "BASELINE.json configs[1]: full fused step, 4096 envs, 1 x B200",
                                 "value": small_m["value"], "unit": UNIT, "ms_per_control_step": small_m["total_ms"] / small_m["ctrl_steps"],
                                 "e2e": small_m["e2e"], "roofline_frac_hbm": a4 / peak, "gpu_launches": small_m["launches"]}
        if e2e is not None:
            line["e2e"] = e2e
        if mdp_only is not None:
            line["mdp_only_kernel"] = mdp_only
        if other_tasks is not None:
            line["other_tasks"] = other_tasks
        if world == 1 and not args.no_cpu_baseline:
            cs = args.cpu_sample_steps
            if cs is None:   # bounded sample of the same workload: ~10 s of CPU work
                _, dt0, _ = cpu_path(n_envs, 10, 1)
                cs = int(max(20, min(20000, 10.0 / (dt0 / 10))))
            v, dt, cores = cpu_path(n_envs, cs, 2)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_envs} envs x {cs} control steps, oracle/cpu_port.cpp float32 OpenMP, {dt:.1f} s",
                                    "note": "C++ port of the same control step on all host threads -- NOT the reference's torch / PhysX "
                                            "path (closed); a harder baseline than it"}
            # the reference's OWN torch MDP (no physics) on the host cores -- only where a reference tree exists
            # (ZBOT_REFERENCE_ROOT; never on the stock GPU box): north_star "reference's torch ... CPU path timed beside it"
            try:
                from oracle import ref_torch_bench
                if ref_torch_bench.available():
                    line["cpu_baseline_torch_mdp"] = ref_torch_bench.time_reference_torch_mdp(n_envs, 10, 2)
                else:
                    line["cpu_baseline_torch_mdp"] = {"unavailable": "no reference tree on this box (ZBOT_REFERENCE_ROOT); measured in the "
                                                      "build container: see profiles/r2_notes.md"}
            except Exception as exc:
                line["cpu_baseline_torch_mdp"] = {"error": repr(exc)}
            # BASELINE.json configs[4] comparator: the same 24 x 4096 rollout (CPU port of the env step + the same MLP) on the host cores
            if other_tasks is not None and "ms_per_rollout" in other_tasks.get("ppo_rollout_24x4096", {}):
                try:
                    import bench_rollout
                    import zbot_lab_b200.tasks.zbot6b_direct.walking_v2 as w2
                    sys.path.insert(0, os.path.join(ROOT, "tests"))
                    from fake_stepper import FakeStepper
                    real = w2.NativeStepper
                    w2.NativeStepper = FakeStepper
                    try:
                        torch.set_num_threads(os.cpu_count() or 1)
                        rc = bench_rollout.make(4096, "cpu", False)
                        sec = bench_rollout.time_rollouts(rc, 3)
                    finally:
                        w2.NativeStepper = real
                    r = other_tasks["ppo_rollout_24x4096"]
                    r["cpu_path"] = {"ms_per_rollout": 1e3 * sec, "value": 24 * 4096 / sec, "unit": UNIT, "threads": torch.get_num_threads(),
                                     "what": "CPU port of the env step (oracle/cpu_port.cpp) + the same actor / critic MLP in CPU torch"}
                    r["speedup_vs_cpu_path"] = sec / (r["ms_per_rollout"] * 1e-3)
                except Exception as exc:
                    other_tasks["ppo_rollout_24x4096"]["cpu_path"] = {"error": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
