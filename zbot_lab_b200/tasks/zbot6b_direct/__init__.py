"""Registers ``zbot-6b-walking-v2`` and ``zbot-6b-walking-v4`` with the same id / kwargs keys as the reference
(``/root/reference/source/zbot/zbot/tasks/zbot6b_direct/__init__.py:41-49``)."""
from ...compat import gym_registry as gym
from .walking_v2 import ZbotDirectEnvV2
from .walking_v2_cfg import PPORunnerCfgV2, ZbotDirectEnvCfgV2
from .walking_v4 import Zbot6SEnvV4
from .walking_v4_cfg import Zbot6SEnvV4Cfg, Zbot6SEnvV4PPOCfg

gym.register(
    id="zbot-6b-walking-v2",
    entry_point="zbot_lab_b200.tasks.zbot6b_direct:ZbotDirectEnvV2",
    disable_env_checker=True,
    kwargs={
        "env_cfg_entry_point": ZbotDirectEnvCfgV2,
        "rsl_rl_cfg_entry_point": f"{__name__}.walking_v2_cfg:PPORunnerCfgV2",
    },
)

# reference: zbot6b_direct/__init__.py:91-99
gym.register(
    id="zbot-6b-walking-v4",
    entry_point="zbot_lab_b200.tasks.zbot6b_direct:Zbot6SEnvV4",
    disable_env_checker=True,
    kwargs={
        "env_cfg_entry_point": Zbot6SEnvV4Cfg,
        "rsl_rl_cfg_entry_point": f"{__name__}.walking_v4_cfg:Zbot6SEnvV4PPOCfg",
    },
)

__all__ = ["ZbotDirectEnvV2", "ZbotDirectEnvCfgV2", "PPORunnerCfgV2", "Zbot6SEnvV4", "Zbot6SEnvV4Cfg", "Zbot6SEnvV4PPOCfg"]
