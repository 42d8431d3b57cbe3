"""shim of ``isaaclab.app.AppLauncher`` (train.py:13,37,48): there is no simulator app to launch --
the environment step is a CUDA kernel -- so this only carries the CLI flags and the local rank."""
import argparse
import os


class _App:
    """``simulation_app``: play.py loops ``while simulation_app.is_running()``.  With no GUI to close, the
    loop length is bounded by ``ZBOT_PLAY_STEPS`` (default: run until interrupted)."""

    def __init__(self):
        n = os.environ.get("ZBOT_PLAY_STEPS")
        self._left = int(n) if n else None

    def is_running(self):
        if self._left is None:
            return True
        self._left -= 1
        return self._left >= 0

    def close(self):
        pass


class AppLauncher:
    def __init__(self, launcher_args: argparse.Namespace | dict | None = None, **kwargs):
        self.args = launcher_args
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.global_rank = int(os.environ.get("RANK", "0"))
        self.app = _App()

    @staticmethod
    def add_app_launcher_args(parser: argparse.ArgumentParser):
        g = parser.add_argument_group("app_launcher", description="(shim) Isaac Sim app arguments")
        g.add_argument("--headless", action="store_true", default=False)
        g.add_argument("--livestream", type=int, default=-1)
        g.add_argument("--enable_cameras", action="store_true", default=False)
        g.add_argument("--device", type=str, default=None)
        g.add_argument("--experience", type=str, default="")
        g.add_argument("--kit_args", type=str, default="")
