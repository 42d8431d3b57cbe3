from zbot_lab_b200.compat.gym_registry import load_cfg_from_registry  # noqa: F401
