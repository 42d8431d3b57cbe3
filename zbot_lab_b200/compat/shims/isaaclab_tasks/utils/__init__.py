"""shim of ``isaaclab_tasks.utils`` (train.py:98; zbot/tasks/__init__.py:3)."""
import os
import re

from zbot_lab_b200.compat.gym_registry import load_cfg_from_registry  # noqa: F401


def import_packages(package_name: str, blacklist_pkgs=None):
    return None


def get_checkpoint_path(log_path: str, run_dir: str = ".*", checkpoint: str = ".*", other_dirs=None, sort_alpha=True) -> str:
    runs = sorted(os.path.join(log_path, d) for d in os.listdir(log_path) if re.match(run_dir, d))
    if not runs:
        raise ValueError(f"No runs present in the directory: '{log_path}' match: '{run_dir}'.")
    run_path = runs[-1]
    files = [f for f in os.listdir(run_path) if re.match(checkpoint, f)]
    if not files:
        raise ValueError(f"No checkpoints in the directory: '{run_path}' match '{checkpoint}'.")
    files.sort(key=lambda m: f"{m:0>15}")
    return os.path.join(run_path, files[-1])
