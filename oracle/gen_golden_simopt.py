"""TEST INFRASTRUCTURE ONLY -- golden vectors for the simulation-optimisation objective, produced by the
UNMODIFIED reference (simopt/pybullet.py:130-225: ObjectiveFunctionPyBullet.evaluate_once / loss_function /
set_parameters), imported from /root/reference with the stand-ins of oracle/shim/.

    python oracle/gen_golden_simopt.py          # writes tests/golden_collector/simopt_hover.npz

The reference fits (thrust-to-weight ratio, motor time constant, latency) to real-flight CSV logs under
data/sim_opt/ -- which are not part of /root/reference.  The "flight log" here is synthetic: the reference's own
DroneHoverBulletEnv flown with near-hover actions under parameters the objective does not know (t2w 1.93,
T 0.11 s), cut into mini-trajectories exactly as simopt/core.py:47-80 does (T = 35, 5 pre-steps, skip 10).
Recorded: the data set, a handful of candidate parameter vectors and evaluate_once()'s loss for every
(candidate, mini-trajectory) pair.  The ring stays 2 sub-steps long: the
engine's ring length is structural.  NOTE set_latency (agents.py:388-404) sizes the ring with int(latency / TIME_STEP)
while the constructor uses int(latency // TIME_STEP): 0.015 s gives 2 sub-steps at construction but 3 after any
set_parameters call (0.015 / 0.005 = 3.0000000000000004).  The candidates use 0.0125 s: 2 sub-steps either way.
"""
import os
import sys


import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get('PHOENIX_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'shim'))
sys.path.insert(0, REFERENCE)
# simopt/plot_utils.py imports matplotlib (absent here) for a plotting helper the objective never calls
import types                                                                 # noqa: E402
_plt = types.ModuleType('matplotlib.pyplot')
_mpl = types.ModuleType('matplotlib')
_mpl.pyplot = _plt
sys.modules.setdefault('matplotlib', _mpl)
sys.modules.setdefault('matplotlib.pyplot', _plt)


from phoenix_drone_simulation.envs.hover import DroneHoverBulletEnv          # noqa: E402
from phoenix_drone_simulation.simopt import pybullet as so                   # noqa: E402


def synthetic_flight(n_steps, seed):
    """xyz, xyz_dot, rpy, rpy_dot rows (the CSV layout of simopt/pybullet.py:143-145) and the actions flown."""
    np.random.seed(seed)
    env = DroneHoverBulletEnv(motor_thrust_noise=0.0)
    env.domain_randomization = -1
    env.observation_noise = -1
    env.enable_reset_distribution = False
    env.drone.update_motor_dynamics(new_motor_time_constant=0.11, new_thrust_to_weight_ratio=1.93)
    env.reset()
    rng = np.random.default_rng(seed)
    rows, acts = [], []
    hover = 2.0 / 1.93 - 1.0
    for t in range(n_steps):
        d = env.drone
        rows.append(np.concatenate([d.xyz, d.xyz_dot, d.rpy, d.rpy_dot]))
        a = (hover + 0.03 * np.sin(0.21 * t + np.arange(4)) + 0.02 * rng.standard_normal(4)).astype(np.float64)
        acts.append(a)
        env.step(a)
    return np.array(rows), np.array(acts)


def slices(obs, acs, T=35, pre=5, skip=10):                       # simopt/core.py:47-80
    o, a, p = [], [], []
    for i in range(pre, obs.shape[0] - T, skip):
        o.append(obs[i:i + T]); a.append(acs[i:i + T]); p.append(acs[i - pre:i])
    return np.array(o), np.array(a), np.array(p)


class _Data:
    mini_trajectory_size = 35


class Objective(so.ObjectiveFunctionHoverTask):
    def _load_real_world_data(self):
        return _Data()

    def _load_simulation(self):
        # simopt/pybullet.py:261-272 verbatim except its last line `_env.seed(seed)`: gymnasium environments have
        # no seed() method any more (the reference's own call fails with the gymnasium it depends on)
        # ... and the motor thrust noise (an Ornstein-Uhlenbeck process fed by the global numpy generator) is
        # switched off through the constructor argument, so that the recorded losses are a function of the data
        # and the candidate alone
        _env = DroneHoverBulletEnv(motor_thrust_noise=0.0)
        _env.domain_randomization = -1
        _env.observation_noise = -1
        np.random.seed(self.seed)
        return _env


def main():
    obs, acs = synthetic_flight(130, 3)
    O, A, P = slices(obs, acs)
    of = Objective(seed=7)
    of.real_data.observations, of.real_data.actions, of.real_data.pre_inputs = O, A, P
    cands = np.array([[1.8, 0.08, 0.0125], [1.93, 0.11, 0.0125], [2.2, 0.05, 0.0125], [1.6, 0.2, 0.0125], [2.05, 0.03, 0.0125]])
    losses = np.zeros((len(cands), len(O)))
    # evaluate_once leaves enable_reset_distribution = False behind (simopt/pybullet.py:157): only the very first
    # call of a process draws a random start for its pre-steps.  One unrecorded call puts the objective into the
    # state every later call sees.
    of.set_parameters(cands[0])
    of.evaluate_once(O[0], A[0], pre_inputs=P[0])
    for k, c in enumerate(cands):
        of.set_parameters(c)
        for m in range(len(O)):
            losses[k, m] = of.evaluate_once(O[m], A[m], pre_inputs=P[m])
    out = os.path.join(ROOT, 'tests', 'golden_collector', 'simopt_hover.npz')
    np.savez_compressed(out, observations=O, actions=A, pre_inputs=P, candidates=cands, losses=losses)
    print('mini-trajectories', O.shape, 'losses per candidate', losses.mean(1))

    # Latency as a parameter (set_parameters -> set_latency, agents.py:388-404): none (below one 5 ms sub-step), 1, 2,
    # 3 (0.015 / 0.005 = 3.0000000000000004), 4, 7 and 10 sub-steps of delay
    cands = np.array([[1.8, 0.08, 0.0], [1.8, 0.08, 0.004], [1.93, 0.11, 0.005], [1.93, 0.11, 0.0125], [2.2, 0.05, 0.015],
                      [2.0, 0.1, 0.021], [1.7, 0.06, 0.0351], [2.1, 0.09, 0.05]])
    losses = np.zeros((len(cands), len(O)))
    rings = []
    for k, c in enumerate(cands):
        of.set_parameters(c)
        d = of.sim_env.drone
        rings.append(d.buf_size if d.use_latency else 0)
        for m in range(len(O)):
            losses[k, m] = of.evaluate_once(O[m], A[m], pre_inputs=P[m])
    out = os.path.join(ROOT, 'tests', 'golden_collector', 'simopt_hover_latency.npz')
    np.savez_compressed(out, observations=O, actions=A, pre_inputs=P, candidates=cands, losses=losses, ring_lengths=np.array(rings))
    print('latency candidates: ring lengths', rings, 'losses per candidate', losses.mean(1))


if __name__ == '__main__':
    main()
