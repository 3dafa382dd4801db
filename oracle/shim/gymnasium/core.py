"""TEST INFRASTRUCTURE ONLY (see gymnasium/__init__.py)."""
from typing import Any, TypeVar

ObsType = TypeVar('ObsType')
ActType = TypeVar('ActType')
RenderFrame = TypeVar('RenderFrame')


class Env:
    metadata: dict = {}
    observation_space: Any = None
    action_space: Any = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)

    def step(self, action):
        obs, r, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, r, terminated, truncated, info
