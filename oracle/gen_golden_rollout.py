"""TEST INFRASTRUCTURE ONLY -- a golden rollout recorded from the UNMODIFIED reference's IWPGAlgorithm.roll_out
(algs/iwpg/iwpg.py:350-385: act with ActorCritic.step, env.step, Buffer.store, finish_path at episode ends /
the epoch cut, env.reset), imported from /root/reference with the stand-ins of oracle/shim/.

    python oracle/gen_golden_rollout.py         # writes tests/golden_collector/rollout_hover_simple.npz

roll_out is called as the reference wrote it (the unbound method on a namespace that carries exactly the
attributes it touches: env, ac, buf, logger, local_steps_per_epoch, max_ep_len), with the reference's own
ActorCritic (core.py:313-411), Buffer (core.py:413-557) and environment.  Recorded: the environment's
standardised random draws per reset / step (as oracle/gen_golden.py), the networks, the normalisers, the
exploration noise and every array the Buffer holds after the epoch.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import gen_golden as gg                                                    # noqa: E402  (sets up the shim + reference paths)

import gymnasium as gym                                                     # noqa: E402
from phoenix_drone_simulation.algs import core                              # noqa: E402
from phoenix_drone_simulation.algs.iwpg.iwpg import IWPGAlgorithm           # noqa: E402


class PhaseEnv:
    """Delegates to the reference env and opens a new tape phase before every reset / step."""

    def __init__(self, env, rec):
        self.env, self.rec = env, rec
        self.reset_tapes, self.step_tapes, self.reset_after, self.n_steps = [], [], [], 0

    def reset(self, **kw):
        self.reset_tapes.append(self.rec.new_phase())
        self.reset_after.append(self.n_steps - 1)
        return self.env.reset(**kw)

    def step(self, a):
        self.step_tapes.append(self.rec.new_phase())
        self.n_steps += 1
        return self.env.step(a)


class Log:
    def __init__(self):
        self.ep_ret, self.ep_len = [], []

    def store(self, **kw):
        if 'EpRet' in kw:
            self.ep_ret.append(kw['EpRet']); self.ep_len.append(kw['EpLen'])


def main(env_id='DroneHoverSimpleEnv-v0', T=160, seed=5):
    np.random.seed(seed)
    torch.manual_seed(seed)
    rec = gg.Recorder()
    with rec:
        init_tape = rec.new_phase()
        env = gym.make(env_id)
        ac_kwargs = {'pi': {'hidden_sizes': (50, 50), 'activation': 'relu'}, 'val': {'hidden_sizes': (64, 64), 'activation': 'tanh'}}
        u = env.unwrapped
        ac = core.ActorCritic('mlp', u.observation_space, u.action_space, ac_kwargs,
                              use_standardized_obs=True, use_scaled_rewards=True)
        # a normaliser that has seen data, a return scale, and less exploration noise than the initial 0.5 (so that
        # episodes last: near-hover actions)
        with torch.no_grad():
            for name, p in ac.named_parameters():
                if name.endswith('4.weight'):
                    p.mul_(0.05)
            ac.pi.net[4].bias.copy_(torch.full((4,), float(env.unwrapped.drone.HOVER_ACTION)))
        ac.obs_oms.mean.data = torch.as_tensor(np.random.default_rng(1).normal(0, 0.2, u.observation_space.shape[0]).astype(np.float32))
        ac.obs_oms.std.data = torch.as_tensor(np.random.default_rng(2).uniform(0.5, 1.5, u.observation_space.shape[0]).astype(np.float32))
        ac.ret_oms.std.data = torch.as_tensor(np.array([7.5], dtype=np.float32))
        ac.pi.set_log_std(0.1)
        buf = core.Buffer(ac, u.observation_space.shape, u.action_space.shape, T, 0.99, 0.95, 'gae', True, True, False)
        penv = PhaseEnv(env, rec)
        stub = types.SimpleNamespace(env=penv, ac=ac, buf=buf, logger=Log(), local_steps_per_epoch=T,
                                     max_ep_len=env._max_episode_steps)
        IWPGAlgorithm.roll_out(stub)

    def pad(rows):
        w = max((len(r) for r in rows), default=0)
        out = np.zeros((len(rows), w))
        for i, r in enumerate(rows):
            out[i, :len(r)] = r
        return out

    sd = {k: v.detach().numpy() for k, v in ac.state_dict().items()}
    out = dict(env_id=env_id, T=T, init_tape=np.array(init_tape), reset_tape=pad(penv.reset_tapes), step_tape=pad(penv.step_tapes),
               reset_after=np.array(penv.reset_after), obs=buf.obs_buf, act=buf.act_buf, rew=buf.rew_buf, val=buf.val_buf,
               logp=buf.logp_buf, adv=buf.adv_buf, target_v=buf.target_val_buf, discounted_ret=buf.discounted_ret_buf,
               ep_ret=np.array(stub.logger.ep_ret), ep_len=np.array(stub.logger.ep_len), **{'sd.' + k: v for k, v in sd.items()})
    path = os.path.join(ROOT, 'tests', 'golden_collector', 'rollout_hover_simple.npz')
    np.savez_compressed(path, **out)
    print(env_id, 'T', T, 'episodes finished', len(stub.logger.ep_len), 'lengths', stub.logger.ep_len, 'resets', len(penv.reset_tapes))


if __name__ == '__main__':
    main()
