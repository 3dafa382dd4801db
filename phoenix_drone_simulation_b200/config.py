"""Environment configuration: the reference's constructor kwargs -> PdxConfig.

Names and defaults follow the reference (paths relative to its phoenix_drone_simulation/):
  base.py:26-48 (DroneBaseEnv kwargs), hover.py:7-24 / circle.py:7-35 / takeoff.py:13-41
  (task defaults), hover.py:253-282 etc. (the concrete Simple/Bullet classes),
  agents.py:142-206 (derived constants), sensors.py:18-33 (noise model),
  envs/assets/cf21x_sys_eq.urdf:10,16-17 and cf21x_bullet.urdf:12,18,30,59-136 (physical
  constants; only the numbers are taken, meshes/rooms are out of scope).
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib

# --- physical constants of the two drone models (URDF <properties>, <mass>, <inertia>) -----
DRONE_MODELS = {
    'cf21x_sys_eq': dict(M=0.027, L=0.0397, T2W=2.25, IXX=1.7e-5, IYY=1.7e-5, IZZ=2.9e-5, PROP_Z=0.0),
    'cf21x_bullet': dict(M=0.030, L=0.0397, T2W=1.8, IXX=1.33e-5, IYY=1.33e-5, IZZ=2.64e-5, PROP_Z=0.0108),
}
KF = 3.16e-10
GND_EFF_COEFF = 11.36859
PROP_RADIUS = 2.31348e-2
DRAG_COEFF_XY, DRAG_COEFF_Z = 9.1785e-7, 10.311e-7
PROP_XY = [(0.028, -0.028), (-0.028, -0.028), (-0.028, 0.028), (0.028, 0.028)]
COLLISION_HALF_HEIGHT = 0.0125

# env id -> (task, physics string, drone model, sim_freq, default aggregate_phy_steps)
ENV_IDS = {
    'DroneHoverSimpleEnv-v0': ('hover', 'SimplePhysics', 'cf21x_sys_eq', 100, 1),
    'DroneHoverBulletEnv-v0': ('hover', 'PyBulletPhysics', 'cf21x_bullet', 200, 2),
    'DroneCircleSimpleEnv-v0': ('circle', 'SimplePhysics', 'cf21x_sys_eq', 100, 1),
    'DroneCircleBulletEnv-v0': ('circle', 'PyBulletPhysics', 'cf21x_bullet', 200, 2),
    'DroneTakeOffSimpleEnv-v0': ('takeoff', 'SimplePhysics', 'cf21x_sys_eq', 100, 1),
    'DroneTakeOffBulletEnv-v0': ('takeoff', 'PyBulletPhysics', 'cf21x_bullet', 200, 2),
}
MAX_EPISODE_STEPS = 500          # __init__.py:11

TASK_DEFAULTS = {
    'hover': dict(penalty_action=1e-4, penalty_angle=0., penalty_spin=1e-4, penalty_terminal=100.,
                  penalty_velocity=0., ARP=0.),
    'circle': dict(penalty_action=1e-4, penalty_angle=0., penalty_spin=1e-3, penalty_terminal=100.,
                   penalty_velocity=1e-4, ARP=1e-3),
    'takeoff': dict(penalty_action=1e-4, penalty_angle=0., penalty_spin=1e-4, penalty_terminal=100.,
                    penalty_velocity=0., ARP=0.),
}


@dataclass
class EnvConfig:
    """Frozen description of one environment flavour (same kwarg names as the reference)."""
    env_id: str
    domain_randomization: float = 0.10
    observation_noise: float = 1
    observation_history_size: int = 2
    aggregate_phy_steps: int = None
    control_mode: str = 'PWM'
    latency: float = 0.015
    motor_time_constant: float = 0.080
    motor_thrust_noise: float = 0.05
    enable_reset_distribution: bool = True
    target_pos: tuple = (0.0, 0.0, 1.0)
    penalty_action: float = None
    penalty_angle: float = None
    penalty_spin: float = None
    penalty_terminal: float = None
    penalty_velocity: float = None
    observation_frequency: int = 100
    max_episode_steps: int = MAX_EPISODE_STEPS
    # extensions (not in the reference)
    use_ground_effect: bool = False
    reset_on_nonfinite: bool = False
    auto_reset: bool = True
    lin_damping: float = 0.04
    ang_damping: float = 0.04
    # None = the agent's default (agents.py:165,196); the reference's debug scripts switch them off by assigning
    # to drone.use_latency / drone.use_motor_dynamics after construction
    use_latency: object = None
    use_motor_dynamics: object = None
    render_mode: object = None
    debug: bool = False
    extra: dict = field(default_factory=dict)

    def __post_init__(self):
        if self.env_id not in ENV_IDS:
            raise KeyError(f'unknown env id {self.env_id!r}; known: {sorted(ENV_IDS)}')
        if self.control_mode not in _lib.PDX_CTRL:
            raise NotImplementedError('control_mode %r: expected one of %s' % (self.control_mode, sorted(_lib.PDX_CTRL)))
        if self.render_mode not in (None, 'rgb_array'):
            raise NotImplementedError('rendering (PyBullet GUI) is not provided')
        self.task, self.physics, self.drone_model, self.sim_freq, default_agg = ENV_IDS[self.env_id]
        if self.aggregate_phy_steps is None or (self.task == 'takeoff' and self.physics == 'SimplePhysics'):
            self.aggregate_phy_steps = default_agg          # takeoff.py:224 hard-codes 1
        for k, v in TASK_DEFAULTS[self.task].items():
            if k != 'ARP' and getattr(self, k) is None:
                setattr(self, k, v)
        self.ARP = TASK_DEFAULTS[self.task]['ARP']

    # ------------------------------------------------------------------------------------
    def to_pdx(self, dtype_code=_lib.PDX_DTYPE_F32, rng_mode=_lib.PDX_RNG_PHILOX):
        mdl = DRONE_MODELS[self.drone_model]
        bullet = self.physics == 'PyBulletPhysics'
        c = _lib.PdxConfig()
        c.task = _lib.PDX_TASK[self.task]
        c.physics = _lib.PDX_PHYSICS[self.physics]
        c.dtype = dtype_code
        c.rng_mode = rng_mode
        c.control_mode = _lib.PDX_CTRL[self.control_mode]                      # agents.py:71-78
        c.observation_noise = 1 if self.observation_noise > 0 else 0
        c.history = int(self.observation_history_size)
        c.agg = int(self.aggregate_phy_steps)
        c.obs_rate = int(self.sim_freq // self.observation_frequency)           # base.py:108
        time_step = 1. / self.sim_freq                                          # base.py:98
        c.use_latency = 1 if (bullet and self.latency >= time_step) else 0      # agents.py:165
        if self.use_latency is not None and bullet:
            c.use_latency = 1 if (self.use_latency and self.latency >= time_step) else 0
        c.buf_size = int(max(1, int(self.latency // time_step)))                # agents.py:180
        c.use_motor_dynamics = 1 if bullet else 0
        if self.use_motor_dynamics is not None and bullet:
            c.use_motor_dynamics = 1 if self.use_motor_dynamics else 0
        c.reset_distribution = 1 if self.enable_reset_distribution else 0
        c.ground_effect = 1 if self.use_ground_effect else 0
        c.max_episode_steps = int(self.max_episode_steps)
        c.reset_on_nonfinite = 1 if self.reset_on_nonfinite else 0
        c.auto_reset = 1 if self.auto_reset else 0
        c.domain_randomization = float(self.domain_randomization)
        c.time_step = time_step
        c.sensor_dt = 1 / self.sim_freq
        G = 9.81                                                                # agents.py:145
        c.mass = mdl['M']
        c.inertia[:] = [mdl['IXX'], mdl['IYY'], mdl['IZZ']]
        c.arm = mdl['L']
        c.gravity = G
        c.thrust2weight = mdl['T2W']
        gravity_force = G * mdl['M']
        c.max_thrust = gravity_force * mdl['T2W'] / 4                           # agents.py:148-149
        c.k_mass_dr = 0.028                                                     # agents.py:224
        c.ftf0, c.ftf1 = 1.56e-5, 5.96e-3                                       # agents.py:142-143
        c.hover_x = float(np.sqrt(1 / mdl['T2W']))                              # agents.py:152
        c.hover_action = 2 * 1 / mdl['T2W'] - 1                                 # agents.py:153
        c.motor_time_constant = float(self.motor_time_constant)
        c.ou_theta = 0.15                                                       # envs/utils.py:88
        c.ou_sigma = 0.2 * self.motor_thrust_noise                              # agents.py:206
        c.lpf_ratio = (1 / self.sim_freq) / (2 / self.sim_freq)                 # base.py:109-110
        c.pos_norm_std, c.pos_unif_range, c.vel_norm_std = 0.002, 0.001, 0.01   # sensors.py:20-23
        c.quat_norm_std = np.pi * 0.1 / 180                                     # sensors.py:24
        c.quat_unif_range = np.pi * 0.05 / 180
        sdt = 1 / self.sim_freq
        sg = 0.000175 / (sdt ** 0.5)                                            # sensors.py:124-128
        c.gyro_sigma_b = (-(sg ** 2) * (1000. / 2) * (math.exp(-2 * sdt / 1000.) - 1)) ** 0.5
        c.gyro_pi = math.exp(-sdt / 1000.)
        c.gyro_random_walk = 0.0105
        c.gyro_turn_on = np.pi * 5 / 180
        c.penalty_action = self.penalty_action
        c.penalty_angle = self.penalty_angle
        c.penalty_spin = self.penalty_spin
        c.penalty_terminal = self.penalty_terminal
        c.penalty_velocity = self.penalty_velocity
        c.action_rate_penalty = self.ARP
        c.target_pos[:] = [float(np.float32(v)) for v in self.target_pos]      # float32 array, hover.py:14
        z0 = 0.0125 if self.task == 'takeoff' else 1.0
        c.init_xyz[:] = [0.0, 0.0, float(np.float32(z0))]                       # hover.py:44, takeoff.py:51
        c.drag_coeff[:] = [DRAG_COEFF_XY, DRAG_COEFF_XY, DRAG_COEFF_Z]
        for i, (x, y) in enumerate(PROP_XY):
            c.prop_xy[i][0], c.prop_xy[i][1] = x, y
        c.prop_z = mdl['PROP_Z']
        c.gnd_eff_coeff, c.prop_radius = GND_EFF_COEFF, PROP_RADIUS
        max_rpm = np.sqrt((mdl['T2W'] * gravity_force) / (4 * c.max_thrust))    # agents.py:155-156
        c.gnd_eff_h_clip = float(0.25 * PROP_RADIUS * np.sqrt(
            (15 * max_rpm ** 2 * KF * GND_EFF_COEFF) / c.max_thrust))
        c.lin_damping, c.ang_damping = self.lin_damping, self.ang_damping
        c.ground_z = COLLISION_HALF_HEIGHT if bullet else 0.0
        _lib.check(_lib.load().pdx_config_finalize(c))
        return c

    @property
    def hover_action(self):
        return 2 * 1 / DRONE_MODELS[self.drone_model]['T2W'] - 1
