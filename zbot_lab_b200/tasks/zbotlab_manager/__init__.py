"""Registers the four manager-based ids with the reference's ids and kwargs keys
(``/root/reference/source/zbot/zbot/tasks/zbotlab_manager/config/zbot6b_manager/__init__.py:14-52``):
``zbot-6b-walking-m-v0`` / ``-m-play-v0`` (flat) and ``zbot-6b-walking-m-rough-v0`` / ``-m-rough-play-v0`` (generated height
field + terrain curriculum, ``zbot_lab_b200/terrain.py``)."""
from ...compat import gym_registry as gym
from .env_cfg import (Zbot6BFlatEnvCfg, Zbot6BFlatEnvCfg_PLAY, Zbot6BFlatPPORunnerCfg, Zbot6BRoughEnvCfg,
                      Zbot6BRoughEnvCfg_PLAY, Zbot6BRoughPPORunnerCfg)
from .manager_env import ManagerBasedRLEnv

gym.register(
    id="zbot-6b-walking-m-v0",
    entry_point="zbot_lab_b200.tasks.zbotlab_manager:ManagerBasedRLEnv",
    disable_env_checker=True,
    kwargs={
        "env_cfg_entry_point": f"{__name__}.env_cfg:Zbot6BFlatEnvCfg",
        "rsl_rl_cfg_entry_point": f"{__name__}.env_cfg:Zbot6BFlatPPORunnerCfg",
    },
)
gym.register(
    id="zbot-6b-walking-m-play-v0",
    entry_point="zbot_lab_b200.tasks.zbotlab_manager:ManagerBasedRLEnv",
    disable_env_checker=True,
    kwargs={
        "env_cfg_entry_point": f"{__name__}.env_cfg:Zbot6BFlatEnvCfg_PLAY",
        "rsl_rl_cfg_entry_point": f"{__name__}.env_cfg:Zbot6BFlatPPORunnerCfg",
    },
)

gym.register(
    id="zbot-6b-walking-m-rough-v0",
    entry_point="zbot_lab_b200.tasks.zbotlab_manager:ManagerBasedRLEnv",
    disable_env_checker=True,
    kwargs={
        "env_cfg_entry_point": f"{__name__}.env_cfg:Zbot6BRoughEnvCfg",
        "rsl_rl_cfg_entry_point": f"{__name__}.env_cfg:Zbot6BRoughPPORunnerCfg",
    },
)
gym.register(
    id="zbot-6b-walking-m-rough-play-v0",
    entry_point="zbot_lab_b200.tasks.zbotlab_manager:ManagerBasedRLEnv",
    disable_env_checker=True,
    kwargs={
        "env_cfg_entry_point": f"{__name__}.env_cfg:Zbot6BRoughEnvCfg_PLAY",
        "rsl_rl_cfg_entry_point": f"{__name__}.env_cfg:Zbot6BRoughPPORunnerCfg",
    },
)

__all__ = ["ManagerBasedRLEnv", "Zbot6BFlatEnvCfg", "Zbot6BFlatEnvCfg_PLAY", "Zbot6BFlatPPORunnerCfg", "Zbot6BRoughEnvCfg",
           "Zbot6BRoughEnvCfg_PLAY", "Zbot6BRoughPPORunnerCfg"]
