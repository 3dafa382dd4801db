"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py

It imports `/root/reference/phoenix_drone_simulation` verbatim, with the stand-ins of
oracle/shim/ on sys.path for the packages the image lacks (pybullet, gymnasium, mpi4py),
drives each env id with a fixed action sequence and records
  * every observation / reward / terminated flag / cost the reference returns,
  * the drone state after every step (xyz, rpy, xyz_dot, rpy_dot),
  * the *standardised random draws* the reference consumed, split per phase
    (constructor, each reset, each step) -- the "tape".
The tape is captured by replacing np.random.{normal,uniform,randn,randint} with
wrappers that draw the underlying standard normal / unit uniform from the very same
global generator and apply `loc + scale*z` / `low + (high-low)*u` themselves, which is
the arithmetic numpy's legacy generator performs; `check_recorder()` asserts that a
recorded run is bit-identical to an unrecorded one.

The fixtures travel to the GPU box; /root/reference does not.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get('PHOENIX_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'shim'))
sys.path.insert(0, REFERENCE)

import gymnasium as gym                     # noqa: E402  (the stand-in)
import phoenix_drone_simulation             # noqa: E402,F401  (the reference, registers ids)


class Recorder:
    """Context manager that records standardised draws of the global numpy generator."""

    def __init__(self):
        self.cur = []
        self._orig = {}

    def new_phase(self):
        self.cur = []
        return self.cur

    def _log(self, x):
        self.cur.extend(np.atleast_1d(np.asarray(x, dtype=np.float64)).ravel().tolist())

    def __enter__(self):
        r = np.random
        self._orig = dict(normal=r.normal, uniform=r.uniform, randn=r.randn, randint=r.randint)
        std_normal, unit = r.standard_normal, r.random_sample

        def normal(loc=0.0, scale=1.0, size=None):
            z = std_normal(size)
            self._log(z)
            return loc + scale * z

        def uniform(low=0.0, high=1.0, size=None):
            u = unit(size)
            self._log(u)
            return low + (np.asarray(high) - np.asarray(low)) * u if size is not None \
                else low + (high - low) * u

        def randn(*shape):
            z = std_normal(shape if shape else None)
            self._log(z)
            return z

        def randint(low, high=None, size=None):
            v = self._orig['randint'](low, high, size)
            self._log(v)
            return v

        r.normal, r.uniform, r.randn, r.randint = normal, uniform, randn, randint
        return self

    def __exit__(self, *exc):
        for k, v in self._orig.items():
            setattr(np.random, k, v)


def make_actions(kind, T, hover_action, seed):
    rng = np.random.default_rng(seed)
    if kind == 'uniform':
        return rng.uniform(-1, 1, (T, 4)).astype(np.float32)
    if kind == 'hover':
        return (hover_action + 0.01 * rng.standard_normal((T, 4))).astype(np.float32)
    if kind == 'mixed':       # a third random, the rest near hover; includes out-of-range values
        a = rng.uniform(-1.3, 1.3, (T, 4)).astype(np.float32)
        a[T // 3:] = (hover_action + 0.05 * rng.standard_normal((T - T // 3, 4))).astype(np.float32)
        return a
    if kind == 'takeoff':
        return (-0.1 + 0.1 * rng.standard_normal((T, 4))).astype(np.float32)
    if kind == 'pid':         # thrust command near hover, small rate / angle commands, some outliers
        a = (0.3 * rng.standard_normal((T, 4))).astype(np.float32)
        a[:, 0] = (hover_action + 0.2 * rng.standard_normal(T)).astype(np.float32)
        a[::17] = rng.uniform(-1.3, 1.3, (len(a[::17]), 4)).astype(np.float32)
        return a
    raise ValueError(kind)


def run_reference(env_id, kwargs, actions_kind, T, seed, record=True, max_episode_steps=500):
    np.random.seed(seed)
    rec = Recorder()
    ctx = rec if record else _Null()
    with ctx:
        init_tape = rec.new_phase()
        # `use_ground_effect` is not a constructor argument of the reference: its physics classes carry the flag
        # (physics.py:18,24) but nothing ever sets it.  The golden switches it on by hand so that the reference's own
        # calculate_ground_effect (physics.py:27-58) and its use in step_forward (physics.py:117-120) are recorded.
        ref_kwargs = {k: v for k, v in kwargs.items() if k != 'use_ground_effect'}
        env = gym.make(env_id, **ref_kwargs)
        u = env.unwrapped
        if kwargs.get('use_ground_effect'):
            u.physics.use_ground_effect = True
        actions = make_actions(actions_kind, T, u.drone.HOVER_ACTION, seed + 1)
        reset_tapes, step_tapes = [], []
        obs, rew, term, cost, state = [], [], [], [], []
        reset_obs, reset_after = [], []

        def snap():
            d = u.drone
            return np.concatenate([d.xyz, d.rpy, d.xyz_dot, d.rpy_dot])

        reset_tapes.append(rec.new_phase())
        o, _ = env.reset()
        reset_obs.append(np.asarray(o, dtype=np.float64))
        reset_after.append(-1)
        reset_state = [snap()]
        for t in range(T):
            step_tapes.append(rec.new_phase())
            o, r, terminated, truncated, info = env.step(actions[t])
            obs.append(np.asarray(o, dtype=np.float64))
            rew.append(float(r))
            term.append(bool(terminated))
            cost.append(float(info['cost']))
            state.append(snap())
            if terminated or truncated:
                reset_tapes.append(rec.new_phase())
                o, _ = env.reset()
                reset_obs.append(np.asarray(o, dtype=np.float64))
                reset_after.append(t)
                reset_state.append(snap())

    def pad(rows):
        w = max((len(r) for r in rows), default=0)
        out = np.zeros((len(rows), w))
        for i, r in enumerate(rows):
            out[i, :len(r)] = r
        return out

    return dict(
        env_id=env_id, kwargs=json.dumps(kwargs), seed=seed, actions_kind=actions_kind,
        actions=actions, init_tape=np.array(init_tape, dtype=np.float64),
        reset_tape=pad(reset_tapes), step_tape=pad(step_tapes),
        obs=np.array(obs), rew=np.array(rew), terminated=np.array(term), cost=np.array(cost),
        state=np.array(state), reset_obs=np.array(reset_obs),
        reset_after=np.array(reset_after, dtype=np.int64), reset_state=np.array(reset_state),
        max_episode_steps=max_episode_steps,
    )


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def check_recorder():
    """The recording wrappers must not change a single bit of the reference's outputs."""
    for env_id in ('DroneHoverSimpleEnv-v0', 'DroneCircleBulletEnv-v0'):
        a = run_reference(env_id, {}, 'mixed', 120, 5, record=True)
        b = run_reference(env_id, {}, 'mixed', 120, 5, record=False)
        for k in ('obs', 'rew', 'terminated', 'cost', 'reset_obs'):
            assert np.array_equal(a[k], b[k]), (env_id, k)


DET = dict(observation_noise=0, domain_randomization=-1, motor_thrust_noise=0.0)

CASES = [
    # name, env id, kwargs, actions, T, seed
    # --- BASELINE.json configs[0]: HoverSimple, 1 env, random actions, 1000 steps ---
    ('config1_hover_simple_default', 'DroneHoverSimpleEnv-v0', {}, 'uniform', 1000, 11),
    ('config1_hover_simple_det', 'DroneHoverSimpleEnv-v0', DET, 'uniform', 1000, 12),
    # 500-step near-hover run without reset distribution (drift figure)
    ('hover_simple_nearhover_det', 'DroneHoverSimpleEnv-v0',
     dict(DET, enable_reset_distribution=False), 'hover', 500, 13),
    ('hover_simple_h3', 'DroneHoverSimpleEnv-v0', dict(observation_history_size=3), 'mixed', 200, 14),
    ('hover_simple_h1_noreset', 'DroneHoverSimpleEnv-v0',
     dict(observation_history_size=1, enable_reset_distribution=False), 'mixed', 150, 15),
    ('circle_simple_default', 'DroneCircleSimpleEnv-v0', {}, 'mixed', 300, 21),
    ('circle_simple_det', 'DroneCircleSimpleEnv-v0', DET, 'mixed', 300, 22),
    ('circle_simple_h8', 'DroneCircleSimpleEnv-v0', dict(observation_history_size=8), 'mixed', 150, 23),
    ('takeoff_simple_default', 'DroneTakeOffSimpleEnv-v0', {}, 'takeoff', 600, 31),
    ('takeoff_simple_det', 'DroneTakeOffSimpleEnv-v0', DET, 'takeoff', 600, 32),
    ('hover_bullet_default', 'DroneHoverBulletEnv-v0', {}, 'mixed', 300, 41),
    ('hover_bullet_det', 'DroneHoverBulletEnv-v0', DET, 'mixed', 300, 42),
    ('hover_bullet_h3', 'DroneHoverBulletEnv-v0', dict(observation_history_size=3), 'mixed', 150, 43),
    ('circle_bullet_default', 'DroneCircleBulletEnv-v0', {}, 'mixed', 300, 51),
    ('takeoff_bullet_default', 'DroneTakeOffBulletEnv-v0', {}, 'takeoff', 300, 61),
    # --- ground effect (physics.py:27-58,117-120; the flag is forced on, see run_reference): take-off starts on the
    # ground where the effect is largest; the hover case flies at 1 m where it is ~1e-4 of the thrust ---
    ('takeoff_bullet_groundeffect', 'DroneTakeOffBulletEnv-v0', dict(use_ground_effect=True), 'takeoff', 300, 62),
    ('takeoff_bullet_groundeffect_det', 'DroneTakeOffBulletEnv-v0', dict(DET, use_ground_effect=True), 'takeoff', 300, 63),
    ('hover_bullet_groundeffect', 'DroneHoverBulletEnv-v0', dict(use_ground_effect=True), 'mixed', 150, 64),
    # --- PID control modes (experiments/07 uses agg 4 / 8 with the Bullet ids) ---
    ('hover_simple_attrate', 'DroneHoverSimpleEnv-v0', dict(control_mode='AttitudeRate'), 'pid', 300, 71),
    ('hover_simple_attitude_det', 'DroneHoverSimpleEnv-v0', dict(DET, control_mode='Attitude'), 'pid', 300, 72),
    ('circle_bullet_attrate_agg4', 'DroneCircleBulletEnv-v0',
     dict(control_mode='AttitudeRate', aggregate_phy_steps=4), 'pid', 200, 73),
    ('hover_bullet_attitude_agg8', 'DroneHoverBulletEnv-v0',
     dict(control_mode='Attitude', aggregate_phy_steps=8), 'pid', 150, 74),
]


def main():
    check_recorder()
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]                      # optional: names of the cases to (re)generate
    for name, env_id, kwargs, kind, T, seed in CASES:
        if only and name not in only:
            continue
        g = run_reference(env_id, kwargs, kind, T, seed)
        np.savez_compressed(os.path.join(out_dir, name + '.npz'), **g)
        print(f'{name:34s} T={T} resets={len(g["reset_after"])} obs_dim={g["obs"].shape[1]} '
              f'tape(reset,step)=({g["reset_tape"].shape[1]},{g["step_tape"].shape[1]}) '
              f'init={g["init_tape"].shape[0]}')


if __name__ == '__main__':
    main()
