"""The reference's OWN ``scripts/rsl_rl/train.py`` runs UNCHANGED on top of this repo's import shims
(zbot_lab_b200/compat/shims) -- build container only (needs /root/reference).  The CUDA stepper is
replaced by a CPU test double (tests/fake_stepper.py) because this box has no GPU; everything else
(registry, cfg classes, env class, rsl_rl wrapper, PPO runner) is the shipped host code."""
import os
import runpy
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SCRIPTS = "/root/reference/scripts/rsl_rl"
SHIMS = os.path.join(ROOT, "zbot_lab_b200", "compat", "shims")


@pytest.fixture
def shimmed(monkeypatch, tmp_path):
    if not os.path.isfile(os.path.join(REF_SCRIPTS, "train.py")):
        pytest.skip("reference tree not present (GPU box)")
    import zbot_lab_b200.tasks.zbot6b_direct.walking_v2 as w2
    from fake_stepper import FakeStepper
    monkeypatch.setattr(w2, "NativeStepper", FakeStepper)
    monkeypatch.syspath_prepend(SHIMS)
    monkeypatch.syspath_prepend(REF_SCRIPTS)
    monkeypatch.chdir(tmp_path)
    for m in [k for k in sys.modules if k.split(".")[0] in ("isaaclab", "isaaclab_rl", "isaaclab_tasks", "gymnasium",
                                                            "omni", "rsl_rl", "zbot", "cli_args")]:
        monkeypatch.delitem(sys.modules, m)
    return tmp_path


def test_reference_train_py_runs_unchanged(shimmed, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["train.py", "--task", "zbot-6b-walking-v2", "--num_envs", "16",
                                      "--max_iterations", "2", "--headless", "--device", "cpu", "--seed", "7",
                                      "agent.num_steps_per_env=6", "agent.device=cpu"])
    runpy.run_path(os.path.join(REF_SCRIPTS, "train.py"), run_name="__main__")
    runs = os.listdir(shimmed / "logs" / "rsl_rl" / "zbot_6b_flat_direct_v2")
    assert len(runs) == 1
    run = shimmed / "logs" / "rsl_rl" / "zbot_6b_flat_direct_v2" / runs[0]
    assert (run / "params" / "env.yaml").exists() and (run / "params" / "agent.pkl").exists()
    assert (run / "model_2.pt").exists()
    import json
    recs = [json.loads(l) for l in open(run / "progress.jsonl")]
    assert len(recs) == 2 and "Episode_Reward/step_length" in recs[-1]


def test_env_surface_on_cpu_double(shimmed):
    """Host logic of ZbotDirectEnvV2 + wrapper on the CPU double: shapes, dtypes, counters, log keys."""
    import torch
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "env_cfg_entry_point")
    cfg.scene.num_envs = 8
    cfg.sim.device = "cpu"
    env = gym.make("zbot-6b-walking-v2", cfg=cfg, render_mode=None)
    w = RslRlVecEnvWrapper(env)
    obs = w.get_observations()["policy"]
    assert obs.shape == (8, 23)
    w.episode_length_buf = torch.full((8,), 997, dtype=torch.int64)
    _, _, dones, ex = w.step(torch.zeros(8, 6))
    assert dones.sum() == 0
    _, _, dones, ex = w.step(torch.zeros(8, 6))
    assert dones.all() and ex["time_outs"].all()                 # 997 + 2 = 999 -> truncation
    assert float(ex["log"]["Episode_Termination/time_out"]) == 8.0
    assert int(env.episode_length_buf.max()) < 1000               # all envs reset -> randint spread


def test_reference_play_py_runs_unchanged(shimmed, monkeypatch):
    """train (2 iterations) -> the reference's own play.py loads the checkpoint, exports the policy and
    steps the env for a bounded number of iterations -- both scripts unmodified."""
    monkeypatch.setattr(sys, "argv", ["train.py", "--task", "zbot-6b-walking-v2", "--num_envs", "8",
                                      "--max_iterations", "1", "--headless", "--device", "cpu",
                                      "agent.num_steps_per_env=4", "agent.device=cpu"])
    runpy.run_path(os.path.join(REF_SCRIPTS, "train.py"), run_name="__main__")
    for m in [k for k in sys.modules if k.split(".")[0] in ("cli_args",)]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.setenv("ZBOT_PLAY_STEPS", "5")
    monkeypatch.setattr(sys, "argv", ["play.py", "--task", "zbot-6b-walking-v2", "--num_envs", "4", "--headless",
                                      "--device", "cpu", "agent.device=cpu"])
    runpy.run_path(os.path.join(REF_SCRIPTS, "play.py"), run_name="__main__")
    root = shimmed / "logs" / "rsl_rl" / "zbot_6b_flat_direct_v2"
    run = root / os.listdir(root)[0]
    assert (run / "exported" / "policy.pt").exists()
    import torch
    pol = torch.jit.load(str(run / "exported" / "policy.pt"))
    assert pol(torch.zeros(3, 23)).shape == (3, 6)


def test_reference_train_py_runs_unchanged_on_the_manager_task(shimmed, monkeypatch):
    """``--task zbot-6b-walking-m-v0``: the reference's own train.py resolves the id (registered by ``import zbot.tasks``)
    to ``isaaclab.envs:ManagerBasedRLEnv``-shaped env + ``Zbot6BFlatEnvCfg`` + the flat PPO cfg, trains two iterations
    and writes its run directory (CPU double of the stepper; the cfg -> term-table compilation, the startup friction
    event, reset, logging and the wrapper are the shipped host code)."""
    import zbot_lab_b200.tasks.zbotlab_manager.manager_env as me
    from fake_stepper import FakeMStepper
    monkeypatch.setattr(me, "NativeStepper", FakeMStepper)
    monkeypatch.setattr(sys, "argv", ["train.py", "--task", "zbot-6b-walking-m-v0", "--num_envs", "16",
                                      "--max_iterations", "2", "--headless", "--device", "cpu", "--seed", "7",
                                      "agent.num_steps_per_env=6", "agent.device=cpu"])
    runpy.run_path(os.path.join(REF_SCRIPTS, "train.py"), run_name="__main__")
    root = shimmed / "logs" / "rsl_rl" / "zbot_6b_flat_mana_v1"
    runs = os.listdir(root)
    assert len(runs) == 1
    run = root / runs[0]
    assert (run / "params" / "env.yaml").exists() and (run / "model_2.pt").exists()
    import json
    recs = [json.loads(l) for l in open(run / "progress.jsonl")]
    assert len(recs) == 2 and "Episode_Reward/foot_step_length" in recs[-1] and "Curriculum/lin_vel_cmd_levels" in recs[-1]


def test_reference_train_py_runs_unchanged_on_the_rough_manager_task(shimmed, monkeypatch):
    """``--task zbot-6b-walking-m-rough-v0`` (config/zbot6b_manager/__init__.py:34-42): generated height field, all 16
    weighted RewTerms of RewardsCfg incl. undesired_contacts, terrain_levels + lin_vel_cmd_levels curricula, the rough agent
    cfg (512-256-128).  CPU double of the stepper; terrain generation, placement, logging are the shipped host code."""
    import zbot_lab_b200.tasks.zbotlab_manager.manager_env as me
    from fake_stepper import FakeMStepper
    monkeypatch.setattr(me, "NativeStepper", FakeMStepper)
    monkeypatch.setattr(sys, "argv", ["train.py", "--task", "zbot-6b-walking-m-rough-v0", "--num_envs", "16",
                                      "--max_iterations", "1", "--headless", "--device", "cpu", "--seed", "7",
                                      "agent.num_steps_per_env=6", "agent.device=cpu"])
    runpy.run_path(os.path.join(REF_SCRIPTS, "train.py"), run_name="__main__")
    root = shimmed / "logs" / "rsl_rl" / "zbot_6b_rough_mana_v1"
    run = root / os.listdir(root)[0]
    assert (run / "params" / "env.yaml").exists() and (run / "model_1.pt").exists()
    import json
    rec = [json.loads(l) for l in open(run / "progress.jsonl")][-1]
    assert "Episode_Reward/undesired_contacts" in rec and "Curriculum/terrain_levels" in rec and 0.0 <= rec["Curriculum/terrain_levels"] <= 5.0


def test_rough_env_places_robots_on_their_tiles(shimmed, monkeypatch):
    """The rough play cfg on the CPU double: 5 x 5 tiles, every env standing on its (level, type) tile's spawn point (the
    base 0.2545 m above the tile's origin), no contact termination in the first steps on any tile family."""
    import torch
    import zbot_lab_b200.tasks.zbotlab_manager.manager_env as me
    from fake_stepper import FakeMStepper
    monkeypatch.setattr(me, "NativeStepper", FakeMStepper)
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-m-rough-play-v0", "env_cfg_entry_point")
    cfg.scene.num_envs, cfg.sim.device, cfg.seed = 25, "cpu", 5
    cfg.terminations.feet_close = None
    cfg.terminations.base_height = None      # root_height_below_minimum is a WORLD height: an inverted pyramid spawns below 0.2 m
    env = gym.make("zbot-6b-walking-m-rough-play-v0", cfg=cfg, render_mode=None)
    t = env._terrain_gen
    assert t.rows == 5 and t.cols == 5 and env._terrain.env_origins.shape == (25, 3)
    lv = env.terrain_levels.numpy()
    ty = env._stepper.state.get("p_delta")[:, 4].numpy().astype(int)
    assert len(set(ty)) == 5 and lv.max() <= 4 and np.allclose(env._terrain.env_origins.numpy(), t.origins[lv, ty])
    env.reset()
    for _ in range(15):
        obs, rew, term, trunc, ex = env.step(torch.zeros(25, 6))
        assert not term.any() and torch.isfinite(rew).all()
    pos, _, _ = env._stepper.articulation_view()
    rough = np.array([t.col_kind[c] == "random_rough" for c in ty])           # no flat platform there: the feet stand on bumps
    dz = (pos[:, 6, 2] - 0.2545).abs().numpy()
    assert dz[~rough].max() < 5e-3 and dz[rough].max() < 0.08 and rough.sum() == 5


def test_manager_env_surface_on_cpu_double(shimmed, monkeypatch):
    """Host logic of ManagerBasedRLEnv + wrapper on the CPU double: 25-wide group, time-out at max_episode_length,
    log keys, command inside the ranges, friction buckets, PLAY cfg."""
    import torch
    import zbot_lab_b200.tasks.zbotlab_manager.manager_env as me
    from fake_stepper import FakeMStepper
    monkeypatch.setattr(me, "NativeStepper", FakeMStepper)
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-m-play-v0", "env_cfg_entry_point")
    cfg.scene.num_envs = 8
    cfg.sim.device = "cpu"
    cfg.seed = 3
    cfg.terminations.feet_close = None          # the init stance sits on that term's margin; keep this test about time-outs
    env = gym.make("zbot-6b-walking-m-play-v0", cfg=cfg, render_mode=None)
    w = RslRlVecEnvWrapper(env)
    obs = w.get_observations()["policy"]
    assert obs.shape == (8, 25) and float(obs[:, 4].abs().max()) <= 0.3 + 1e-6 and float(obs[:, 7:19].abs().max()) < 1e-5
    assert float(env.friction.min()) >= 0.3 and float(env.friction.max()) <= 1.0
    w.episode_length_buf = torch.full((8,), 998, dtype=torch.int64)
    _, _, dones, ex = w.step(torch.zeros(8, 6))
    assert dones.sum() == 0
    obs2, _, dones, ex = w.step(torch.zeros(8, 6))
    assert dones.all() and ex["time_outs"].all()                 # 998 + 2 = 1000 = max_episode_length (mdp.time_out)
    assert float(ex["log"]["Episode_Termination/time_out"]) == 8.0 and float(ex["log"]["Episode_Termination/base_height"]) == 0.0
    assert "Episode_Termination/feet_close" not in ex["log"] and "Episode_Reward/track_lin_vel_xy_exp" in ex["log"]
    assert int(env.episode_length_buf.max()) == 0 and float(obs2["policy"][:, 19:].abs().max()) == 0
    assert env.command.shape == (8, 3)


def test_reference_play_py_runs_unchanged_on_the_manager_task(shimmed, monkeypatch):
    """train one iteration of ``zbot-6b-walking-m-v0`` -> the reference's own play.py with the PLAY id loads the checkpoint,
    exports the 25-input policy and steps the 64-env PLAY cfg (no observation corruption, commands from the limit ranges)."""
    import zbot_lab_b200.tasks.zbotlab_manager.manager_env as me
    from fake_stepper import FakeMStepper
    monkeypatch.setattr(me, "NativeStepper", FakeMStepper)
    monkeypatch.setattr(sys, "argv", ["train.py", "--task", "zbot-6b-walking-m-v0", "--num_envs", "8",
                                      "--max_iterations", "1", "--headless", "--device", "cpu",
                                      "agent.num_steps_per_env=4", "agent.device=cpu"])
    runpy.run_path(os.path.join(REF_SCRIPTS, "train.py"), run_name="__main__")
    for m in [k for k in sys.modules if k.split(".")[0] in ("cli_args",)]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.setenv("ZBOT_PLAY_STEPS", "5")
    monkeypatch.setattr(sys, "argv", ["play.py", "--task", "zbot-6b-walking-m-play-v0", "--num_envs", "4", "--headless",
                                      "--device", "cpu", "agent.device=cpu"])
    runpy.run_path(os.path.join(REF_SCRIPTS, "play.py"), run_name="__main__")
    root = shimmed / "logs" / "rsl_rl" / "zbot_6b_flat_mana_v1"
    run = root / os.listdir(root)[0]
    assert (run / "exported" / "policy.pt").exists()
    import torch
    pol = torch.jit.load(str(run / "exported" / "policy.pt"))
    assert pol(torch.zeros(3, 25)).shape == (3, 6)
