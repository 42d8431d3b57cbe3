"""Minimal stand-in for ``isaaclab.utils.configclass``: attribute-style cfg objects with
``replace`` / ``to_dict`` / ``copy`` so the reference's cfg idioms work
(``env_cfg.scene.num_envs = ...``, ``agent_cfg.to_dict()``; scripts/rsl_rl/train.py:114-155,185)."""
from __future__ import annotations

import copy


class Cfg:
    """Class attributes are defaults; instances deep-copy mutable defaults so that mutating
    one instance never leaks into the class (cf. the reference's in-place ``reward_scales``
    scaling)."""

    def __init__(self, **kwargs):
        for klass in reversed(type(self).__mro__):
            for k, v in vars(klass).items():
                if k.startswith("_") or callable(v) or isinstance(v, (property, staticmethod, classmethod)):
                    continue
                setattr(self, k, copy.deepcopy(v))
        for k, v in kwargs.items():
            setattr(self, k, v)

    def replace(self, **kwargs):
        new = copy.deepcopy(self)
        for k, v in kwargs.items():
            setattr(new, k, v)
        return new

    def copy(self):
        return copy.deepcopy(self)

    def to_dict(self) -> dict:
        out = {}
        for k, v in vars(self).items():
            if k.startswith("_"):
                continue
            out[k] = v.to_dict() if isinstance(v, Cfg) else copy.deepcopy(v)
        return out

    def from_dict(self, d: dict):
        for k, v in d.items():
            cur = getattr(self, k, None)
            if isinstance(cur, Cfg) and isinstance(v, dict):
                cur.from_dict(v)
            else:
                setattr(self, k, v)

    def __repr__(self):
        return f"{type(self).__name__}({self.to_dict()})"


def configclass(cls):
    """Decorator form, for code written against ``@configclass``."""
    if not issubclass(cls, Cfg):
        cls = type(cls.__name__, (cls, Cfg), dict(vars(cls)))
    return cls
