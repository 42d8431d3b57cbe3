"""Registers ``zbot-6s-snake-v0`` with the same id / kwargs keys as the reference
(``/root/reference/source/zbot/zbot/tasks/zbot6_direct/__init__.py:25-33``)."""
from ...compat import gym_registry as gym
from .snake_v0 import ZbotDirectEnvV0
from .snake_v0_cfg import PPORunnerCfgV1, ZbotDirectEnvCfgV0

gym.register(
    id="zbot-6s-snake-v0",
    entry_point="zbot_lab_b200.tasks.zbot6_direct:ZbotDirectEnvV0",
    disable_env_checker=True,
    kwargs={
        "env_cfg_entry_point": ZbotDirectEnvCfgV0,
        "rsl_rl_cfg_entry_point": f"{__name__}.snake_v0_cfg:PPORunnerCfgV1",
    },
)

__all__ = ["ZbotDirectEnvV0", "ZbotDirectEnvCfgV0", "PPORunnerCfgV1"]
