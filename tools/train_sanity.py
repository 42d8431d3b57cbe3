"""Learning sanity of the three tasks on the B200 step: PPO for a few hundred iterations, first vs last 20 iterations.
   python tools/train_sanity.py [iters]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zbot_lab_b200.tasks  # noqa: E402,F401
from zbot_lab_b200.compat import gym_registry as gym  # noqa: E402
from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper  # noqa: E402
from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
tasks = sys.argv[2:] or ["zbot-6b-walking-v2", "zbot-6b-walking-v4", "zbot-6s-snake-v0", "zbot-6b-walking-m-v0"]
out = {}
for task in tasks:
    cfg = gym.load_cfg_from_registry(task, "env_cfg_entry_point")
    cfg.scene.num_envs, cfg.sim.device, cfg.seed = 4096, "cuda:0", 42
    if hasattr(cfg, "events") and hasattr(cfg.events, "my_curric"):
        cfg.events.my_curric = False          # fixed reward table for the comparison
    if hasattr(cfg, "curriculum"):
        cfg.curriculum.lin_vel_cmd_levels = None
    env = RslRlVecEnvWrapper(gym.make(task, cfg=cfg, render_mode=None))
    acfg = gym.load_cfg_from_registry(task, "rsl_rl_cfg_entry_point").to_dict()
    r = OnPolicyRunner(env, acfg, log_dir=None, device="cuda:0")
    h = r.learn(iters, init_at_random_ep_len=True)
    keys = [k for k in h[-1] if k.startswith("Episode_Termination")]
    m = lambda rows, k: sum(x.get(k, 0.0) for x in rows) / len(rows)
    out[task] = {"iterations": iters, "fps_last": h[-1]["fps"],
                 "mean_step_reward_first20": m(h[:20], "mean_step_reward"), "mean_step_reward_last20": m(h[-20:], "mean_step_reward"),
                 **{k + "_first20": m(h[:20], k) for k in keys}, **{k + "_last20": m(h[-20:], k) for k in keys}}
    env.close()
print(json.dumps(out, indent=1))
