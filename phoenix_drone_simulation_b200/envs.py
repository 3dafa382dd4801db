"""Single-environment drop-in: the reference's six env ids with the gymnasium 5-tuple API.

Every env here is an N=1 view over the batched CUDA engine (vec_env.VecEnv); `reset` and
`step` return numpy arrays / Python scalars exactly like the reference
(envs/base.py:382-475):  obs float64 ndarray [D], reward float, terminated bool,
truncated bool, info dict with 'cost'.  Ids and `max_episode_steps=500` follow
phoenix_drone_simulation/__init__.py:8-50.  If `gymnasium` is importable the ids are
registered there as well (so `gymnasium.make(id, **kwargs)` keeps working); otherwise the
local `make()` below is the registry.

Not provided: the PyBullet handle `bc`, `render()` (GUI), LIDAR sensors, PID control modes.
"""
import numpy as np
import torch

from . import lib as _lib
from .config import ENV_IDS, MAX_EPISODE_STEPS, DRONE_MODELS, EnvConfig
from .vec_env import VecEnv

try:                                    # optional
    import gymnasium as _gym
except Exception:                       # pragma: no cover - gymnasium is not in this image
    _gym = None


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (used when gymnasium is absent)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


def _box(low, high):
    if _gym is not None:
        return _gym.spaces.Box(low, high, dtype=np.float32)
    return Box(low, high)


class _DroneView:
    """`env.drone`: read/write access to the agent state (agents.py:52-56, 152-153, 185)."""

    def __init__(self, env):
        self._env = env
        mdl = DRONE_MODELS[env.config.drone_model]
        self.HOVER_ACTION = 2 * 1 / mdl['T2W'] - 1
        self.HOVER_X = float(np.sqrt(1 / mdl['T2W']))
        self.M, self.L, self.THRUST2WEIGHT_RATIO = mdl['M'], mdl['L'], mdl['T2W']
        self.act_dim = 4
        self.use_latency = bool(env._vec.pdx.use_latency)

    def _get(self, name):
        return self._env._vec.get_state(name)[0].double().cpu().numpy()

    @property
    def xyz(self):
        return self._get('xyz')

    @xyz.setter
    def xyz(self, v):
        self._env._vec.set_state('xyz', np.asarray(v, dtype=np.float64))

    @property
    def xyz_dot(self):
        return self._get('vel')

    @xyz_dot.setter
    def xyz_dot(self, v):
        self._env._vec.set_state('vel', np.asarray(v, dtype=np.float64))

    @property
    def quaternion(self):
        if self._env.config.physics == 'PyBulletPhysics':
            return self._get('quat')
        r, p, y = self.rpy / 2.0
        q = np.array([np.sin(r) * np.cos(p) * np.cos(y) - np.cos(r) * np.sin(p) * np.sin(y),
                      np.cos(r) * np.sin(p) * np.cos(y) + np.sin(r) * np.cos(p) * np.sin(y),
                      np.cos(r) * np.cos(p) * np.sin(y) - np.sin(r) * np.sin(p) * np.cos(y),
                      np.cos(r) * np.cos(p) * np.cos(y) + np.sin(r) * np.sin(p) * np.sin(y)])
        return q / np.linalg.norm(q)

    @property
    def rpy(self):
        if self._env.config.physics == 'SimplePhysics':
            return self._get('rpy')
        x, y, z, w = self._get('quat')
        sarg = -2.0 * (x * z - w * y)
        if sarg <= -0.99999:
            return np.array([0.0, -0.5 * np.pi, 2 * np.arctan2(x, -y)])
        if sarg >= 0.99999:
            return np.array([0.0, 0.5 * np.pi, 2 * np.arctan2(-x, y)])
        return np.array([np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), np.arcsin(sarg),
                         np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)])

    @rpy.setter
    def rpy(self, v):
        if self._env.config.physics != 'SimplePhysics':
            raise NotImplementedError('set the quaternion state for Bullet ids')
        self._env._vec.set_state('rpy', np.asarray(v, dtype=np.float64))

    @property
    def rpy_dot(self):
        if self._env.config.physics == 'SimplePhysics':
            return self._get('omega')
        raise NotImplementedError('body rates of Bullet ids: use get_state("omega_world")')

    @rpy_dot.setter
    def rpy_dot(self, v):
        self._env._vec.set_state('omega', np.asarray(v, dtype=np.float64))

    @property
    def last_action(self):
        return self._get('last_action')

    @property
    def x(self):
        return self._get('motor_x')


class DroneEnv:
    """gymnasium-style single environment backed by the CUDA engine (N = 1)."""
    metadata = {'render.modes': ['rgb_array']}

    def __init__(self, env_id, device='cuda', dtype=torch.float64, seed=None, **kwargs):
        kwargs = dict(kwargs)
        kwargs['auto_reset'] = False             # the caller resets, as in the reference
        self.config = EnvConfig(env_id, **kwargs)
        self.env_id = env_id
        self._seed = int(np.random.SeedSequence().entropy % (2 ** 63)) if seed is None else int(seed)
        self._vec = VecEnv(env_id, 1, device=device, dtype=dtype, seed=self._seed, config=self.config)
        self._max_episode_steps = self.config.max_episode_steps
        self.observation_history_size = self.config.observation_history_size
        self.domain_randomization = self.config.domain_randomization
        self.observation_noise = self.config.observation_noise
        self.enable_reset_distribution = self.config.enable_reset_distribution
        self.aggregate_phy_steps = self.config.aggregate_phy_steps
        self.SIM_FREQ = self.config.sim_freq
        self.TIME_STEP = 1. / self.SIM_FREQ
        self.render_mode = self.config.render_mode
        obs_dim = self._vec.obs_dim
        o_lim = 1000 * np.ones((obs_dim,), dtype=np.float32)       # base.py:147-150
        a_lim = np.ones((4,), dtype=np.float32)
        self.observation_space = _box(-o_lim, o_lim)
        self.action_space = _box(-a_lim, a_lim)
        self.drone = _DroneView(self)
        self._needs_reset = True

    # gymnasium plumbing -----------------------------------------------------------------------
    @property
    def unwrapped(self):
        return self

    @property
    def time_step(self):
        return float(self._vec.get_state('dt')[0, 0])

    @property
    def iteration(self):
        return int(self._vec.get_state('ep_length')[0, 0]) * self.aggregate_phy_steps

    def close(self):
        pass

    def render(self):
        if self.render_mode == 'rgb_array':
            return np.array([])
        raise NotImplementedError('the PyBullet GUI is not part of the B200 engine')

    def reset(self, *, seed=None, options=None):
        """Like the reference (base.py:385) `seed` does not re-seed the environment's noise
        stream; it only seeds the action space sampler."""
        if seed is not None and hasattr(self.action_space, 'seed'):
            self.action_space.seed(seed)
        obs = self._vec.reset()
        self._needs_reset = False
        return obs[0].double().cpu().numpy(), {}

    def step(self, action):
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, 4), device=self._vec.device)
        obs, rew, term, trunc, info = self._vec.step(a)
        out = torch.cat([obs[0].double(), rew.double(), info['cost'].double(),
                         term.double(), trunc.double()]).cpu().numpy()        # one D2H copy
        d = self._vec.obs_dim
        return (out[:d], float(out[d]), bool(out[d + 2]), bool(out[d + 3]), {'cost': float(out[d + 1])})


def _make_cls(env_id):
    name = env_id.split('-')[0]

    class _Env(DroneEnv):
        def __init__(self, **kwargs):
            super().__init__(env_id, **kwargs)
    _Env.__name__ = _Env.__qualname__ = name
    return _Env


DroneHoverSimpleEnv = _make_cls('DroneHoverSimpleEnv-v0')
DroneHoverBulletEnv = _make_cls('DroneHoverBulletEnv-v0')
DroneCircleSimpleEnv = _make_cls('DroneCircleSimpleEnv-v0')
DroneCircleBulletEnv = _make_cls('DroneCircleBulletEnv-v0')
DroneTakeOffSimpleEnv = _make_cls('DroneTakeOffSimpleEnv-v0')
DroneTakeOffBulletEnv = _make_cls('DroneTakeOffBulletEnv-v0')

registry = {env_id: globals()[env_id.split('-')[0]] for env_id in ENV_IDS}


def make(env_id, **kwargs):
    """Local equivalent of `gymnasium.make(env_id, **kwargs)` for the six Drone ids.  The time
    limit (500 steps) is enforced inside the kernel, so no TimeLimit wrapper is needed."""
    if env_id not in registry:
        raise KeyError(f'unknown env id {env_id!r}; known: {sorted(registry)}')
    return registry[env_id](**kwargs)


def register_with_gymnasium():
    if _gym is None:
        return False
    for env_id in ENV_IDS:
        if env_id not in _gym.envs.registry:
            _gym.register(id=env_id, entry_point=f'{__name__}:{env_id.split("-")[0]}',
                          max_episode_steps=MAX_EPISODE_STEPS)
    return True
