"""TEST INFRASTRUCTURE ONLY -- golden vectors for the attribute surface the reference's own debug script uses
(debug/compare_system_equations_with_PyBullet.py:13-64), produced by the UNMODIFIED reference imported from
/root/reference with the stand-ins of oracle/shim/.

    python oracle/gen_golden_debug_compare.py      # writes tests/golden_collector/debug_compare.npz

The script builds DroneHoverSimpleEnv and DroneHoverBulletEnv, switches domain randomisation, observation noise, the
reset distribution, the latency ring and the motor lag off BY ASSIGNING TO ATTRIBUTES AFTER CONSTRUCTION, then flies
50 steps of one fixed action and reads drone.xyz / rpy / rpy_dot / quaternion before every step.  This generator does
the same through the same attributes and records what they returned, plus the two reset observations.

One deviation from the script: both envs are built with motor_thrust_noise=0.0.  The script leaves the motor
thrust noise (an Ornstein-Uhlenbeck process on the global numpy generator, agents.py:280) running, so its curves differ
from run to run; a fixture needs a deterministic trajectory and the engine draws from Philox, not from numpy.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get('PHOENIX_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'shim'))
sys.path.insert(0, REFERENCE)

from phoenix_drone_simulation.envs.hover import DroneHoverBulletEnv, DroneHoverSimpleEnv     # noqa: E402

N_STEPS = 50


def switch_everything_off(env):
    env.unwrapped.domain_randomization = -1
    env.unwrapped.observation_noise = -1
    env.enable_reset_distribution = False
    env.drone.USE_LATENCY = False
    env.drone.use_latency = False
    env.drone.use_motor_dynamics = False


def fly(env):
    reset_obs, _ = env.reset()
    out = {k: [] for k in ('xyz', 'rpy', 'rpy_dot', 'quaternion', 'xyz_dot', 'obs', 'reward')}
    action = 0.01 * np.ones(4)
    action[3] = 0.5
    for _ in range(N_STEPS):
        d = env.drone
        out['xyz'].append(np.array(d.xyz)); out['rpy'].append(np.array(d.rpy))
        out['rpy_dot'].append(np.array(d.rpy_dot)); out['quaternion'].append(np.array(d.quaternion))
        out['xyz_dot'].append(np.array(d.xyz_dot))
        o, r, term, trunc, info = env.step(action)
        out['obs'].append(np.array(o)); out['reward'].append(r)
    res = {k: np.array(v, dtype=np.float64) for k, v in out.items()}
    res['reset_obs'] = np.array(reset_obs, dtype=np.float64)
    res['action'] = action
    return res


def main():
    np.random.seed(0)
    data = {}
    for tag, cls in (('bullet', DroneHoverBulletEnv), ('simple', DroneHoverSimpleEnv)):
        env = cls(motor_thrust_noise=0.0)
        switch_everything_off(env)
        for k, v in fly(env).items():
            data[f'{tag}_{k}'] = v
        env.close()
    out = os.path.join(ROOT, 'tests', 'golden_collector', 'debug_compare.npz')
    np.savez_compressed(out, **data)
    print('wrote', out, {k: v.shape for k, v in data.items()})
    print('bullet z', data['bullet_xyz'][[0, 10, 49], 2], 'simple z', data['simple_xyz'][[0, 10, 49], 2])
    print('bullet rpy_dot[49]', data['bullet_rpy_dot'][49], 'simple', data['simple_rpy_dot'][49])


if __name__ == '__main__':
    main()
