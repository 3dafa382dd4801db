"""TEST INFRASTRUCTURE ONLY (see gymnasium/__init__.py)."""
import importlib
from dataclasses import dataclass
from typing import Optional

registry = {}


@dataclass
class EnvSpec:
    id: str
    entry_point: str
    max_episode_steps: Optional[int] = None


def register(id, entry_point, max_episode_steps=None, **kwargs):
    registry[id] = EnvSpec(id, entry_point, max_episode_steps)


def make(id, **kwargs):
    from ..core import TimeLimit
    spec = registry[id]
    mod_name, cls_name = spec.entry_point.split(':')
    cls = getattr(importlib.import_module(mod_name), cls_name)
    env = cls(**kwargs)
    if spec.max_episode_steps is not None:
        env = TimeLimit(env, spec.max_episode_steps)
    return env
