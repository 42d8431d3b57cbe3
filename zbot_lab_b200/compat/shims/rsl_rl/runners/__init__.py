"""shim of ``rsl_rl.runners`` (train.py:81)."""
from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner  # noqa: F401


class DistillationRunner:
    def __init__(self, *a, **k):
        raise NotImplementedError("distillation is not part of the zbot-6b-walking-v2 path")
