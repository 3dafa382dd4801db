"""Single-environment drop-in: the reference's six env ids with the gymnasium 5-tuple API.

Every env here is an N=1 view over the batched CUDA engine (vec_env.VecEnv); `reset` and
`step` return numpy arrays / Python scalars exactly like the reference
(envs/base.py:382-475):  obs float64 ndarray [D], reward float, terminated bool,
truncated bool, info dict with 'cost'.  Ids and `max_episode_steps=500` follow
phoenix_drone_simulation/__init__.py:8-50.  If `gymnasium` is importable the ids are
registered there as well (so `gymnasium.make(id, **kwargs)` keeps working); otherwise the
local `make()` below is the registry.

Attribute surface kept for the reference's own scripts (debug/compare_system_equations_with_PyBullet.py:13-64,
simopt/pybullet.py:130-160, utils/evaluation.py): `env.unwrapped`, assignable `domain_randomization`,
`observation_noise`, `enable_reset_distribution` (the engine is re-built with the new switches: assign before
`reset()`, as those scripts do), `init_xyz / init_rpy / init_quaternion / init_xyz_dot / init_rpy_dot` (the state
`reset()` starts from when the reset distribution is off, hover.py:192-243), `observation_history`, and
`env.drone.{xyz, xyz_dot, rpy, rpy_dot, quaternion, x, y, last_action, control, use_latency, use_motor_dynamics}`.

Not provided: the PyBullet handle `bc`, `render()` (GUI), LIDAR sensors.
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
        self.control = type('Control', (), {'name': env.config.control_mode,
                                            '__repr__': lambda c: f'<{env.config.control_mode} control (on device)>'})()

    # agents.py:165,196: switches the reference's debug scripts assign to after construction
    @property
    def use_latency(self):
        return bool(self._env._vec.pdx.use_latency)

    @use_latency.setter
    def use_latency(self, v):
        self._env._reconfigure(use_latency=bool(v))

    USE_LATENCY = use_latency

    @property
    def use_motor_dynamics(self):
        return bool(self._env._vec.pdx.use_motor_dynamics)

    @use_motor_dynamics.setter
    def use_motor_dynamics(self, v):
        self._env._reconfigure(use_motor_dynamics=bool(v))

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

    def _R(self):
        x, y, z, w = self.quaternion
        s_ = 2.0 / (x * x + y * y + z * z + w * w)
        return np.array([[1 - s_ * (y * y + z * z), s_ * (x * y - w * z), s_ * (x * z + w * y)],
                         [s_ * (x * y + w * z), 1 - s_ * (x * x + z * z), s_ * (y * z - w * x)],
                         [s_ * (x * z - w * y), s_ * (y * z + w * x), 1 - s_ * (x * x + y * y)]])

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
        v = np.asarray(v, dtype=np.float64)
        if self._env.config.physics == 'SimplePhysics':
            self._env._vec.set_state('rpy', v)
        else:                                   # Bullet ids integrate the quaternion (agents.py:443-446)
            self.quaternion = _quat_from_euler(v)

    @quaternion.setter
    def quaternion(self, q):
        q = np.asarray(q, dtype=np.float64)
        if self._env.config.physics == 'PyBulletPhysics':
            self._env._vec.set_state('quat', q / np.linalg.norm(q))
        else:                                   # Simple ids integrate the Euler angles (physics.py:178-179)
            self._env._vec.set_state('rpy', _euler_from_quat(q))

    @property
    def rpy_dot(self):
        """Body rates (agents.py:56,452-453)."""
        if self._env.config.physics == 'SimplePhysics':
            return self._get('omega')
        return self._R().T @ self._get('omega_world')

    @rpy_dot.setter
    def rpy_dot(self, v):
        v = np.asarray(v, dtype=np.float64)
        if self._env.config.physics == 'SimplePhysics':
            self._env._vec.set_state('omega', v)
        else:
            self._env._vec.set_state('omega_world', self._R() @ v)

    @property
    def last_action(self):
        return self._get('last_action')

    @property
    def x(self):
        """Motor state (agents.py:185); zeros for agents without motor dynamics."""
        if self._env.config.physics != 'PyBulletPhysics':
            return np.zeros(4)
        return self._get('motor_x')

    @x.setter
    def x(self, v):
        self._env._vec.set_state('motor_x', np.asarray(v, dtype=np.float64))

    @property
    def y(self):
        """Motor forces of the last sub-step without the thrust noise (agents.py:186,292: y = K n)."""
        if self._env.config.physics != 'PyBulletPhysics' or not self.use_motor_dynamics:
            a = np.clip(self.last_action, -1, 1)
            return self._env._vec.pdx.max_thrust * (30000 + a * 30000) / 60000
        x = self.x
        return self._get('motor_k') * np.clip(x * x, 0, 1)

    @y.setter
    def y(self, v):                            # derived here; the reference overwrites it on the next sub-step
        pass


def _quat_from_euler(rpy):
    r, p, y = np.asarray(rpy, dtype=np.float64) / 2.0
    q = np.array([np.sin(r) * np.cos(p) * np.cos(y) - np.cos(r) * np.sin(p) * np.sin(y),
                  np.cos(r) * np.sin(p) * np.cos(y) + np.sin(r) * np.cos(p) * np.sin(y),
                  np.cos(r) * np.cos(p) * np.sin(y) - np.sin(r) * np.sin(p) * np.cos(y),
                  np.cos(r) * np.cos(p) * np.cos(y) + np.sin(r) * np.sin(p) * np.sin(y)])
    return q / np.linalg.norm(q)


def _euler_from_quat(q):
    x, y, z, w = q
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -0.99999:
        return np.array([0.0, -0.5 * np.pi, 2 * np.arctan2(x, -y)])
    if sarg >= 0.99999:
        return np.array([0.0, 0.5 * np.pi, 2 * np.arctan2(-x, y)])
    return np.array([np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), np.arcsin(sarg),
                     np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)])


class DroneEnv:
    """gymnasium-style single environment backed by the CUDA engine (N = 1)."""
    metadata = {'render.modes': ['rgb_array']}

    def __init__(self, env_id, device='cuda', dtype=torch.float64, seed=None, **kwargs):
        kwargs = dict(kwargs)
        kwargs['auto_reset'] = False             # the caller resets, as in the reference
        self.config = EnvConfig(env_id, **kwargs)
        self.env_id = env_id
        self._seed = int(np.random.SeedSequence().entropy % (2 ** 63)) if seed is None else int(seed)
        self._device, self._dtype = device, dtype
        self._vec = VecEnv(env_id, 1, device=device, dtype=dtype, seed=self._seed, config=self.config)
        self._max_episode_steps = self.config.max_episode_steps
        self.observation_history_size = self.config.observation_history_size
        self.aggregate_phy_steps = self.config.aggregate_phy_steps
        # base.py:113-118: the state reset() starts from (before the reset distribution, if enabled)
        self.init_xyz = np.array(self._vec.pdx.init_xyz[:], dtype=np.float32)
        self.init_rpy = np.zeros(3)
        self.init_quaternion = np.array([0.0, 0.0, 0.0, 1.0])
        self.init_xyz_dot = np.zeros(3)
        self.init_rpy_dot = np.zeros(3)
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

    # switches the reference's scripts assign to after construction (debug/compare_system_equations_with_PyBullet.py:
    # 18-30, simopt/pybullet.py:157,263-265).  They are compile-time / layout choices of the engine, so the N = 1
    # engine behind this env is re-built; the scripts assign them before reset().
    def _reconfigure(self, **changes):
        kw = {k: getattr(self.config, k) for k in (
            'domain_randomization', 'observation_noise', 'observation_history_size', 'aggregate_phy_steps', 'control_mode',
            'latency', 'motor_time_constant', 'motor_thrust_noise', 'enable_reset_distribution', 'target_pos', 'penalty_action',
            'penalty_angle', 'penalty_spin', 'penalty_terminal', 'penalty_velocity', 'observation_frequency', 'max_episode_steps',
            'use_ground_effect', 'reset_on_nonfinite', 'auto_reset', 'lin_damping', 'ang_damping', 'use_latency',
            'use_motor_dynamics', 'render_mode', 'debug')}
        kw.update(changes)
        self.config = EnvConfig(self.env_id, **kw)
        self._vec = VecEnv(self.env_id, 1, device=self._device, dtype=self._dtype, seed=self._seed, config=self.config)
        o_lim = 1000 * np.ones((self._vec.obs_dim,), dtype=np.float32)
        self.observation_space = _box(-o_lim, o_lim)
        self._needs_reset = True

    domain_randomization = property(lambda self: self.config.domain_randomization,
                                    lambda self, v: self._reconfigure(domain_randomization=v))
    observation_noise = property(lambda self: self.config.observation_noise,
                                 lambda self, v: self._reconfigure(observation_noise=v))
    enable_reset_distribution = property(lambda self: self.config.enable_reset_distribution,
                                         lambda self, v: self._reconfigure(enable_reset_distribution=bool(v)))

    @property
    def observation_history(self):
        """base.py:136: the H most recent compute_observation() results, oldest first."""
        row = self._vec.obs[0].double().cpu().numpy()
        E = self._vec.core_dim + 4
        return [row[j * E:j * E + self._vec.core_dim].copy() for j in range(self.observation_history_size)]

    def seed(self, seed=None):
        """Old-gym API some of the reference's scripts still call (simopt/pybullet.py:271); seeds the action space."""
        if hasattr(self.action_space, 'seed'):
            self.action_space.seed(seed)
        return [seed]

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
        custom = (np.any(np.asarray(self.init_rpy) != 0) or np.any(np.asarray(self.init_xyz_dot) != 0) or
                  np.any(np.asarray(self.init_rpy_dot) != 0) or not np.allclose(self.init_quaternion, [0, 0, 0, 1]) or
                  not np.allclose(self.init_xyz, self._vec.pdx.init_xyz[:]))
        if custom:
            obs = self._reset_from_init_state()
        return obs[0].double().cpu().numpy(), {}

    def _reset_from_init_state(self):
        """reset() from init_xyz / init_quaternion / init_xyz_dot / init_rpy_dot (hover.py:192-243 with the reset
        distribution off; simopt/pybullet.py:147-158).  Supported for noise-free environments without reset
        distribution: the reset observation is then the state itself."""
        if self.config.enable_reset_distribution or self.config.observation_noise > 0 or self.config.task != 'hover':
            raise NotImplementedError('init_* start states are provided for the hover ids with '
                                      'enable_reset_distribution=False and observation_noise<=0 (the simopt use)')
        d, q = self.drone, np.asarray(self.init_quaternion, dtype=np.float64)
        d.xyz = np.asarray(self.init_xyz, dtype=np.float64)
        d.quaternion = q
        d.xyz_dot = np.asarray(self.init_xyz_dot, dtype=np.float64)
        R = d._R()
        ww = R.T @ np.asarray(self.init_rpy_dot, dtype=np.float64)      # hover.py:242: handed to Bullet as the WORLD rate
        if self.config.physics == 'PyBulletPhysics':
            self._vec.set_state('omega_world', ww)
        else:
            self._vec.set_state('omega', R.T @ ww)                      # agents.py:452-453: R^T again (quirk A.6-4)
        # the reset observation of a noise-free env is [xyz, quat, vel, body rates, ...] repeated H times
        row = self._vec.obs[0].clone()
        E, C = self._vec.core_dim + 4, self._vec.core_dim
        core = torch.as_tensor(np.concatenate([d.xyz, d.quaternion, d.xyz_dot, d.rpy_dot]), dtype=row.dtype, device=row.device)
        for j in range(self.observation_history_size):
            row[j * E:j * E + 13] = core
        self._vec.obs[0].copy_(row)
        hf, hn = self._vec._field('hist')
        for s_ in range(self.observation_history_size - 1):             # the history slots the next step shifts from
            words = row[(s_ + 1) * E:(s_ + 2) * E]
            qh = (E + 3) // 4 * 4
            for k in range(E):
                w = hf + s_ * qh + k
                self._vec.state[w // 4, 0, w % 4] = words[k]
        return self._vec.obs

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
    """Registers the six ids with gymnasium (done on import of this module when gymnasium is installed, like the
    reference's phoenix_drone_simulation/__init__.py:8-50), so that `gymnasium.make(id, **kwargs)` keeps working."""
    if _gym is None:
        return False
    for env_id in ENV_IDS:
        if env_id not in _gym.envs.registry:
            _gym.register(id=env_id, entry_point=f'{__name__}:{env_id.split("-")[0]}',
                          max_episode_steps=MAX_EPISODE_STEPS)
    return True


register_with_gymnasium()
