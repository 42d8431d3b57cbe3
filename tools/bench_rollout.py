"""BASELINE.json configs[4]: a PPO rollout of 24 steps x 4096 envs end to end (policy MLP 23->128->128->128->6
+ critic, env step, storage) -- eager launches vs one captured CUDA graph, next to the CPU path (CPU port of
the env step + the same torch MLP on the host cores).   python tools/bench_rollout.py [envs] [iters]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import zbot_lab_b200.tasks  # noqa: E402,F401
from zbot_lab_b200.compat import gym_registry as gym  # noqa: E402
from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper  # noqa: E402
from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner  # noqa: E402


def make(n, device, graph, fused=True):
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "env_cfg_entry_point")
    cfg.scene.num_envs, cfg.sim.device, cfg.seed = n, device, 1
    cfg.check_all_envs_reset = False
    env = RslRlVecEnvWrapper(gym.make("zbot-6b-walking-v2", cfg=cfg, render_mode=None))
    acfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "rsl_rl_cfg_entry_point").to_dict()
    acfg["use_cuda_graph"] = graph
    acfg["fused_policy"] = fused      # zbot_policy_act / zbot_rollout_store instead of the torch act / store (CUDA envs only)
    return OnPolicyRunner(env, acfg, log_dir=None, device=device)


def time_rollouts(r, iters):
    obs = r.env.get_observations()["policy"]
    if r.use_cuda_graph:
        obs = r.capture_rollout(obs)
    for _ in range(3):
        obs, _ = r.replay_rollout() if r.use_cuda_graph else r.collect_rollout(obs)
    if r.device.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        obs, _ = r.replay_rollout() if r.use_cuda_graph else r.collect_rollout(obs)
    if r.device.type == "cuda":
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = {"config": f"PPO rollout 24 steps x {n} envs (policy+critic MLP 3x128 ELU, env step, storage)"}
    for name, graph, fused in (("torch_policy_eager", False, False), ("torch_policy_cuda_graph", True, False),
                               ("fused_policy_eager", False, True), ("fused_policy_cuda_graph", True, True)):
        r = make(n, "cuda:0", graph, fused)
        assert (r._fused is not None) == fused
        s = time_rollouts(r, iters)
        out[name] = {"ms_per_rollout": 1e3 * s, "us_per_step": 1e6 * s / 24, "env_steps_per_s": 24 * n / s}
        r.env.close()
    # the act launch alone, CUDA events, back to back
    r = make(n, "cuda:0", False, True)
    st, b, pol = r._fused, r.buf, r._policy_struct()
    obs = r.env.get_observations()["policy"]
    def act():
        st.policy_act(pol, obs, b["obs"][0], b["act"][0], b["logp"][0], b["val"][0], b["mu"][0], b["sigma"][0])
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(5):
            act()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()          # 50 launches per replay: the Python around one ctypes call is longer than the kernel
    with torch.cuda.graph(g):
        for _ in range(50):
            act()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 500
    flop = 2.0 * n * 2 * (r.num_obs * 128 + 2 * 128 * 128) + 2.0 * n * 128 * (r.num_actions + 1)
    out["zbot_policy_act_kernel"] = {"us": us, "fp32_tflops": flop / us * 1e-6}
    r.env.close()
    # CPU path: CPU port of the env step (tests/fake_stepper.py double) + the same MLP on the host cores
    if "--no-cpu" not in sys.argv:
        import zbot_lab_b200.tasks.zbot6b_direct.walking_v2 as w2
        from fake_stepper import FakeStepper
        w2.NativeStepper = FakeStepper
        torch.set_num_threads(os.cpu_count() or 1)
        r = make(n, "cpu", False)
        s = time_rollouts(r, max(2, iters // 10))
        out["cpu_path"] = {"ms_per_rollout": 1e3 * s, "env_steps_per_s": 24 * n / s, "threads": torch.get_num_threads()}
    print(json.dumps(out))
