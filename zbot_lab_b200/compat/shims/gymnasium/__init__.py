"""shim of the part of ``gymnasium`` the reference touches (registry + ``spaces.flatdim``)."""
from zbot_lab_b200.compat.gym_registry import make, register, registry, spec  # noqa: F401

from . import spaces  # noqa: F401


class Env:
    pass


class wrappers:  # noqa: N801
    class RecordVideo:
        def __init__(self, env, **kwargs):
            raise NotImplementedError("video recording is out of scope of the B200 step")
