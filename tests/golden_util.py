"""Helpers shared by the CPU and GPU parity tests: load a golden fixture (generated from
the unmodified reference by oracle/gen_golden.py) and replay the oracle on its tape."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g['env_id'] = str(g['env_id'])
    g['kwargs'] = json.loads(str(g['kwargs']))
    g['seed'] = int(g['seed'])
    g['name'] = name
    return g


def replay_oracle(g, source=None, numpy_global=False):
    """Run the oracle over the golden's action sequence with the golden's reset protocol
    (reset after `terminated` or after 500 steps).  Returns arrays shaped like the golden's."""
    from oracle.phoenix_oracle import OracleEnv, TapeSource, NumpyGlobalSource
    if numpy_global:
        np.random.seed(g['seed'])
        src = NumpyGlobalSource()
    else:
        src = source or TapeSource(g['reset_tape'], g['step_tape'], g['init_tape'])
    env = OracleEnv(g['env_id'], src, **g['kwargs'])
    obs, rew, term, cost, state, reset_obs, reset_after, reset_state = [], [], [], [], [], [], [], []

    def snap():
        return np.concatenate([env.xyz, env.rpy, env.vel, env.omega])

    o, _ = env.reset()
    reset_obs.append(o)
    reset_after.append(-1)
    reset_state.append(snap())
    ep_len = 0
    for t in range(g['actions'].shape[0]):
        o, r, terminated, _, info = env.step(g['actions'][t])
        ep_len += 1
        obs.append(o)
        rew.append(r)
        term.append(terminated)
        cost.append(info['cost'])
        state.append(snap())
        if terminated or ep_len == int(g['max_episode_steps']):
            o, _ = env.reset()
            reset_obs.append(o)
            reset_after.append(t)
            reset_state.append(snap())
            ep_len = 0
    return dict(obs=np.array(obs), rew=np.array(rew), terminated=np.array(term),
                cost=np.array(cost), state=np.array(state), reset_obs=np.array(reset_obs),
                reset_after=np.array(reset_after), reset_state=np.array(reset_state))
