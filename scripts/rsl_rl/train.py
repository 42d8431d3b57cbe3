#!/usr/bin/env python
"""Train a ZBOT task (``zbot-6b-walking-v2`` / ``-v4`` / ``zbot-6s-snake-v0``) with PPO on the B200 step.

Same flow and flags as the reference's ``scripts/rsl_rl/train.py`` (task lookup in the gym registry
-> cfg overrides from the CLI -> ``gym.make(task, cfg=env_cfg)`` -> ``RslRlVecEnvWrapper`` ->
``OnPolicyRunner(env, agent_cfg.to_dict(), log_dir, device)`` -> ``runner.learn(...,
init_at_random_ep_len=True)``), on the in-repo equivalents of the Isaac Lab / rsl_rl modules that are not
installable here.  No simulator app is launched: the step is a CUDA kernel.

  python scripts/rsl_rl/train.py --task zbot-6b-walking-v2 --num_envs 4096 --max_iterations 100
  torchrun --nproc-per-node 8 scripts/rsl_rl/train.py --task zbot-6b-walking-v2 --distributed
"""
import argparse
import os
import sys
from datetime import datetime

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

parser = argparse.ArgumentParser(description="Train an RL agent (PPO) on the B200 ZBOT step.")
parser.add_argument("--num_envs", type=int, default=None, help="Number of environments to simulate.")
parser.add_argument("--task", type=str, default="zbot-6b-walking-v2", help="Name of the task.")
parser.add_argument("--agent", type=str, default="rsl_rl_cfg_entry_point")
parser.add_argument("--seed", type=int, default=None, help="Seed used for the environment")
parser.add_argument("--max_iterations", type=int, default=None, help="RL Policy training iterations.")
parser.add_argument("--distributed", action="store_true", default=False, help="Run training with multiple GPUs.")
parser.add_argument("--device", type=str, default=None)
parser.add_argument("--experiment_name", type=str, default=None)
parser.add_argument("--run_name", type=str, default=None)
parser.add_argument("--resume", action="store_true", default=False)
parser.add_argument("--checkpoint", type=str, default=None, help="Checkpoint file to resume from.")
parser.add_argument("--log_root", type=str, default="logs/rsl_rl")
parser.add_argument("--load_run", type=str, default=None, help="Name of the run folder to resume from.")
parser.add_argument("--logger", type=str, default=None, choices={"wandb", "tensorboard", "neptune"})
parser.add_argument("--log_project_name", type=str, default=None)
# flags of the reference command lines that have no effect here (no simulator app, no renderer):
# AppLauncher.add_app_launcher_args (train.py:37) and the video options (train.py:20-22, 33)
parser.add_argument("--headless", action="store_true", default=False, help="accepted for compatibility (always headless)")
parser.add_argument("--enable_cameras", action="store_true", default=False, help="accepted for compatibility")
parser.add_argument("--livestream", type=int, default=0, help="accepted for compatibility")
parser.add_argument("--video", action="store_true", default=False, help="accepted; rendering is out of scope (no video is recorded)")
parser.add_argument("--video_length", type=int, default=200)
parser.add_argument("--video_interval", type=int, default=2000)
parser.add_argument("--export_io_descriptors", action="store_true", default=False)
args_cli, hydra_overrides = parser.parse_known_args()   # the reference forwards the rest to Hydra (train.py:44-45)

import torch  # noqa: E402

import zbot_lab_b200.tasks  # noqa: F401,E402  (registers the task ids)
from zbot_lab_b200 import distributed as zdist  # noqa: E402
from zbot_lab_b200.compat import gym_registry as gym  # noqa: E402
from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper  # noqa: E402
from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner  # noqa: E402


def main():
    env_cfg = gym.load_cfg_from_registry(args_cli.task, "env_cfg_entry_point")
    agent_cfg = gym.load_cfg_from_registry(args_cli.task, args_cli.agent)
    # Hydra-style overrides of the reference CLI ("env.scene.num_envs=64 agent.max_iterations=10", train.py:109)
    for ov in hydra_overrides:
        if "=" not in ov or not ov.split(".", 1)[0] in ("env", "agent"):
            raise SystemExit(f"unrecognized argument: {ov}")
        path, val = ov.split("=", 1)
        obj = env_cfg if path.startswith("env.") else agent_cfg
        *parents, leaf = path.split(".")[1:]
        for part in parents:
            obj = obj[part] if isinstance(obj, dict) else getattr(obj, part)
        try:
            val = __import__("ast").literal_eval(val)
        except (ValueError, SyntaxError):
            pass
        if isinstance(obj, dict):
            obj[leaf] = val
        else:
            setattr(obj, leaf, val)
    if args_cli.video:
        print("[WARN] --video: rendering is out of scope of the B200 step; training without recording.")
    if args_cli.seed is not None:
        agent_cfg.seed = args_cli.seed
    if args_cli.experiment_name:
        agent_cfg.experiment_name = args_cli.experiment_name
    if args_cli.run_name:
        agent_cfg.run_name = args_cli.run_name
    if args_cli.num_envs is not None:
        env_cfg.scene.num_envs = args_cli.num_envs
    if args_cli.max_iterations is not None:
        agent_cfg.max_iterations = args_cli.max_iterations
    env_cfg.seed = agent_cfg.seed
    if args_cli.device is not None:
        env_cfg.sim.device = args_cli.device
        agent_cfg.device = args_cli.device
    if args_cli.distributed:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        env_cfg.sim.device = agent_cfg.device = f"cuda:{local_rank}"
        env_cfg.seed = agent_cfg.seed = zdist.rank_seed(agent_cfg.seed, local_rank)   # train.py:130-132

    log_root_path = os.path.abspath(os.path.join(args_cli.log_root, agent_cfg.experiment_name))
    log_dir = datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
    if agent_cfg.run_name:
        log_dir += f"_{agent_cfg.run_name}"
    log_dir = os.path.join(log_root_path, log_dir)
    env_cfg.log_dir = log_dir
    print(f"[INFO] Logging experiment in directory: {log_root_path}")

    env = gym.make(args_cli.task, cfg=env_cfg, render_mode=None)
    env = RslRlVecEnvWrapper(env, clip_actions=agent_cfg.clip_actions)
    runner = OnPolicyRunner(env, agent_cfg.to_dict(), log_dir=log_dir, device=agent_cfg.device)
    runner.add_git_repo_to_log(__file__)
    if args_cli.resume and args_cli.checkpoint:
        print(f"[INFO]: Loading model checkpoint from: {args_cli.checkpoint}")
        runner.load(args_cli.checkpoint)
    os.makedirs(os.path.join(log_dir, "params"), exist_ok=True)
    import json
    with open(os.path.join(log_dir, "params", "env.json"), "w") as f:
        json.dump(env_cfg.to_dict(), f, indent=1, default=str)
    with open(os.path.join(log_dir, "params", "agent.json"), "w") as f:
        json.dump(agent_cfg.to_dict(), f, indent=1, default=str)
    hist = runner.learn(num_learning_iterations=agent_cfg.max_iterations, init_at_random_ep_len=True)
    if hist:
        last = hist[-1]
        print(f"[INFO] iteration {last['iteration']}: fps {last['fps']:.0f}, mean step reward {last['mean_step_reward']:.4f}")
    env.close()


if __name__ == "__main__":
    main()
