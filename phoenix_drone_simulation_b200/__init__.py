"""phoenix_drone_simulation_b200 -- B200-native batched stepping engine for the Crazyflie
environments of phoenix-drone-simulation (Drone{Hover,Circle,TakeOff}{Simple,Bullet}Env-v0).

Public API
  make(env_id, **kwargs)            single-env drop-in (gymnasium 5-tuple), N=1 view
  VecEnv(env_id, num_envs, ...)     lock-step batched environments on one GPU
  RolloutCollector / OnlineMeanStd  device-side PPO rollout collection (rollout.py)
The compute path is CUDA only (libphoenix_b200.so through the C ABI in
include/phoenix_b200.h); importing the package does not need a GPU, using it does.
"""
from .lib import PhoenixB200Error, LIB_PATH  # noqa: F401
from .config import EnvConfig, ENV_IDS, MAX_EPISODE_STEPS  # noqa: F401


try:                            # like the reference's __init__.py:8-50: importing the package registers the ids
    import gymnasium as _gymnasium_present  # noqa: F401
    from . import envs as _envs  # noqa: F401  (registers on import)
except Exception:               # gymnasium is optional: the local make() below is the registry then
    pass


def __getattr__(name):          # torch-dependent modules are imported lazily
    if name in ('VecEnv',):
        from .vec_env import VecEnv
        return VecEnv
    if name in ('make', 'registry', 'DroneEnv', 'register_with_gymnasium'):
        from . import envs
        return getattr(envs, name)
    if name in ('RolloutCollector', 'OnlineMeanStd', 'ActorCritic', 'compute_gae', 'EpisodeStats'):
        from . import rollout
        return getattr(rollout, name)
    raise AttributeError(name)
