"""shim of the names train.py:83-89 imports from ``isaaclab.envs``."""
from zbot_lab_b200.tasks.zbot6b_direct.walking_v2_cfg import ZbotDirectEnvCfgV2 as DirectRLEnvCfg  # noqa: F401


class DirectMARLEnv:  # no multi-agent envs in scope
    pass


class DirectMARLEnvCfg:
    pass


from zbot_lab_b200.tasks.zbotlab_manager.env_cfg import ManagerBasedRLEnvCfg  # noqa: E402,F401
from zbot_lab_b200.tasks.zbotlab_manager.manager_env import ManagerBasedRLEnv  # noqa: E402,F401


def multi_agent_to_single_agent(env):
    return env
