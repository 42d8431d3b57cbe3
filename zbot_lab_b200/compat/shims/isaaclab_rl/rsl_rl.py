"""shim of ``isaaclab_rl.rsl_rl`` (train.py:93)."""
from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper  # noqa: F401
from zbot_lab_b200.tasks.zbot6b_direct.walking_v2_cfg import (  # noqa: F401
    PPORunnerCfgV2 as RslRlOnPolicyRunnerCfg,
    RslRlPpoActorCriticCfg,
    RslRlPpoAlgorithmCfg,
)
from zbot_lab_b200.utils.configclass import Cfg as RslRlBaseRunnerCfg  # noqa: F401
