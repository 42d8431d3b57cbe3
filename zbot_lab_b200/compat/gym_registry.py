"""Task registry with the ``gymnasium`` surface the reference uses (``gym.register`` /
``gym.make`` / ``gym.spec``; zbot6b_direct/__init__.py:41-49, scripts/rsl_rl/train.py:158).
``gymnasium`` is not installed in this image; when it is importable the ids are registered
there as well."""
from __future__ import annotations

import importlib
from dataclasses import dataclass, field


@dataclass
class EnvSpec:
    id: str
    entry_point: object
    disable_env_checker: bool = True
    kwargs: dict = field(default_factory=dict)


registry: dict[str, EnvSpec] = {}


def register(id: str, entry_point, disable_env_checker: bool = True, kwargs: dict | None = None, **_):
    registry[id] = EnvSpec(id, entry_point, disable_env_checker, dict(kwargs or {}))
    try:  # mirror into gymnasium when present
        import gymnasium

        if id not in gymnasium.registry:
            gymnasium.register(id=id, entry_point=entry_point, disable_env_checker=disable_env_checker,
                               kwargs=dict(kwargs or {}))
    except Exception:
        pass


def spec(id: str) -> EnvSpec:
    if id not in registry:
        raise KeyError(f"No registered env with id: {id}. Known: {sorted(registry)}")
    return registry[id]


def _load(entry_point):
    if callable(entry_point):
        return entry_point
    mod, _, name = entry_point.partition(":")
    return getattr(importlib.import_module(mod), name)


def make(id: str, **kwargs):
    s = spec(id)
    kw = {k: v for k, v in s.kwargs.items() if not k.endswith("_entry_point")}
    kw.update(kwargs)
    return _load(s.entry_point)(**kw)


def load_cfg_from_registry(task: str, entry_point_key: str):
    """``isaaclab_tasks.utils.parse_cfg.load_cfg_from_registry``: class object or "module:Name"."""
    ep = spec(task).kwargs[entry_point_key]
    cls = _load(ep) if isinstance(ep, str) else ep
    return cls() if isinstance(cls, type) else cls
