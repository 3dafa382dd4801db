"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the simulation-optimisation objective of the reference
(phoenix_drone_simulation/simopt/pybullet.py) on top of the environment oracle.  Pinned by
tests/golden_collector/simopt_hover.npz (oracle/gen_golden_simopt.py: the unmodified reference).

  set_parameters   simopt/pybullet.py:232-247 -> envs/agents.py:208-224 (update_motor_dynamics)
  evaluate_once    simopt/pybullet.py:130-183
  loss_function    simopt/pybullet.py:195-225
"""
import numpy as np

from oracle.phoenix_oracle import OracleEnv, euler_from_quat, quat_from_euler, rot_from_quat


class _NoDraws:
    """Random source of an environment that must not draw (noise, randomisation and reset distribution off)."""

    def begin_init(self): pass
    def begin_reset(self, e): pass
    def begin_step(self, t): pass

    def _no(self, *a):
        raise AssertionError('the simulation-optimisation environment drew a random number')
    normal = uniform = uniform1 = randint = _no

    def randn(self, n):              # the OU thrust noise still draws (envs/utils.py:104-108); its sigma is 0 here
        return np.zeros(n)


def make_env(env_id='DroneHoverBulletEnv-v0'):
    # _load_simulation (pybullet.py:261-272): noise and domain randomisation off; motor thrust noise off as in the
    # golden generator; the reset distribution is off from the second evaluate_once call on (pybullet.py:157)
    return OracleEnv(env_id, _NoDraws(), domain_randomization=-1, observation_noise=0, motor_thrust_noise=0.0,
                     enable_reset_distribution=False)


def set_parameters(env, params):
    p = np.clip(np.asarray(params, dtype=np.float64), 0, np.inf)
    T = np.clip(p[1], env.dt, np.inf)                     # agents.py:217-218 (T_s = the nominal time step)
    env.A = np.ones(4) * (1 - env.dt / T)
    env.B = np.ones(4) * (env.dt / T)
    env.K = np.ones(4) * (0.028 * env.G * p[0] / 4)       # agents.py:224
    env.set_latency(p[2])                                 # pybullet.py:248 -> agents.py:388-404


def loss_function(obs_sim, obs_real):
    e_rpy = euler_from_quat(obs_sim[3:7]) - obs_real[6:9]
    e_rpy_dot = obs_sim[10:13] - obs_real[9:12]
    e_xyz = 100 * (obs_sim[0:3] - obs_real[0:3])
    e_xyz_dot = 10 * (obs_sim[7:10] - obs_real[3:6])
    err = np.hstack((e_rpy, e_xyz, e_xyz_dot, e_rpy_dot))
    return np.linalg.norm(err, ord=1) + np.linalg.norm(err, ord=2)


def evaluate_once(env, obs, acs, pre_inputs, gamma=0.95):
    z0 = 1.0
    env.init_xyz = np.array([0, 0, z0], dtype=np.float32)
    env.init_quat = quat_from_euler(np.zeros(3))
    env.init_xyz_dot, env.init_rpy_dot = np.zeros(3), np.zeros(3)
    # NOTE the reference does NOT restore the initial state between calls: the pre-steps of call k start from the
    # logged state of call k-1 (init_* stay set).  The caller passes `carry` through `env.simopt_init`.
    if getattr(env, 'simopt_init', None) is not None:
        env.init_xyz, env.init_quat, env.init_xyz_dot, env.init_rpy_dot = env.simopt_init
    env.reset()
    for u in pre_inputs:
        env.step(u)
    x = env.x.copy()
    x0 = obs[0]
    rpy = x0[6:9]
    q = quat_from_euler(rpy)
    R = rot_from_quat(q)
    env.simopt_init = (np.array(x0[:3]), q, np.array(x0[3:6]), R @ x0[9:12])
    env.init_xyz, env.init_quat, env.init_xyz_dot, env.init_rpy_dot = env.simopt_init
    env.reset()
    env.x = x
    errs = []
    for i in range(obs.shape[0] - 1):
        env.step(acs[i])
        sim_obs = env.obs_hist[-1]
        errs.append(gamma ** i * loss_function(sim_obs, obs[i + 1]))
    return float(np.mean(errs))
