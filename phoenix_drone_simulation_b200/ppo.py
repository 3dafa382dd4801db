"""PPO on the batched engine: rollouts by the device-side collector, update in torch.

Mirrors the reference's on-policy loop (paths relative to phoenix_drone_simulation/):
  algs/iwpg/iwpg.py:259-326  learn / learn_one_epoch (roll_out -> update -> log), linear learning
                             rate decay (:208-221), exploration-noise annealing (:272-273)
  algs/iwpg/iwpg.py:398-485  update(): value net `train_v_iterations` x `num_mini_batches`
                             mini-batch steps, policy net `train_pi_iterations` full-batch steps,
                             gradients averaged across ranks (mpi_avg_grads, mpi_tools.py:30-36)
  algs/ppo/ppo.py:22-40      clipped surrogate loss
  algs/ppo/defaults.py:6-19  networks pi 50-50 relu, v 64-64 tanh, gamma 0.99
The env.step hot path, GAE, running statistics and the fused policy step are CUDA kernels of this
repo; the SGD update itself stays torch autograd (out of the hot-path scope,
section 8f-2 "next").  Ranks: one process per GPU, NCCL all-reduce of gradients.
"""
import math
import time

import torch

from .rollout import ActorCritic, RolloutCollector, _world
from .vec_env import VecEnv


class PPO:
    def __init__(self, env_id, num_envs=4096, steps=64, epochs=50, device='cuda', seed=0, dist=None,
                 gamma=0.99, lam=0.95, clip_ratio=0.2, pi_lr=3e-4, vf_lr=1e-3, train_pi_iterations=80,
                 train_v_iterations=5, num_mini_batches=16, entropy_coef=0.0, target_kl=0.01,
                 use_kl_early_stopping=False, use_linear_lr_decay=True, use_exploration_noise_anneal=True,
                 use_standardized_advantages=False, **env_kwargs):
        rank = dist.get_rank() if _world(dist) > 1 else 0
        self.dist, self.epochs, self.epoch = dist, epochs, 0
        self.env = VecEnv(env_id, num_envs, device=device, seed=seed, env_offset=rank * num_envs,
                          keep_final_obs=True, **env_kwargs)
        torch.manual_seed(seed)                              # same initial weights on every rank (sync_params)
        self.ac = ActorCritic(self.env.obs_dim, device=self.env.device, dist=dist, seed=seed + 10000 * rank)
        # The reference resets the env at the start of every roll_out and demands max_ep_len <=
        # local_steps_per_epoch (iwpg.py:217,353).  With a shorter rollout window that would only ever show
        # the first `steps` steps of an episode, so episodes then continue across rollouts (the collector
        # carries obs[T] and bootstraps the cut with V(obs[T]), iwpg.py:376-378).
        self.collector = RolloutCollector(self.env, self.ac, steps, gamma=gamma, lam=lam, dist=dist,
                                          reset_each_rollout=steps >= self.env.max_episode_steps)
        self.clip_ratio, self.entropy_coef, self.target_kl = clip_ratio, entropy_coef, target_kl
        self.train_pi_iterations, self.train_v_iterations = train_pi_iterations, train_v_iterations
        self.num_mini_batches, self.use_kl_early_stopping = num_mini_batches, use_kl_early_stopping
        self.use_exploration_noise_anneal = use_exploration_noise_anneal
        self.use_standardized_advantages = use_standardized_advantages
        self.pi_opt = torch.optim.Adam(self.ac.pi.parameters(), lr=pi_lr)
        self.vf_opt = torch.optim.Adam(self.ac.v.parameters(), lr=vf_lr)
        self.sched = None
        if use_linear_lr_decay:                              # iwpg.py:208-221
            self.sched = torch.optim.lr_scheduler.LambdaLR(self.pi_opt, lambda e: 1 - e / max(1, epochs))
        self.history = []

    # ---------------------------------------------------------------------------------------
    def _avg_grads(self, module):
        P = _world(self.dist)
        if P > 1:
            for p in module.parameters():
                self.dist.all_reduce(p.grad, op=self.dist.ReduceOp.SUM)
                p.grad /= P

    def _log_prob(self, obs_std, act):
        mu = self.ac.pi(obs_std)
        std = torch.exp(self.ac.log_std)
        logp = (-0.5 * ((act - mu) / std) ** 2 - self.ac.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)
        return mu, std, logp

    def _loss_pi(self, obs, act, adv, logp_old):            # ppo.py:22-40
        mu, std, logp = self._log_prob(obs, act)
        ratio = torch.exp(logp - logp_old)
        clip_adv = adv * torch.clamp(ratio, 1 - self.clip_ratio, 1 + self.clip_ratio)
        loss = -torch.min(ratio * adv, clip_adv).mean()
        if self.entropy_coef:
            # ppo.py:32: `loss_pi -= entropy_coef * dist.entropy().mean()` -- the mean over the batch AND the
            # action dimensions of the per-dimension Gaussian entropy
            loss = loss - self.entropy_coef * (0.5 + 0.5 * math.log(2 * math.pi) + self.ac.log_std).mean()
        return loss, mu, std

    def update(self, data):
        T, N = data['rew'].shape
        obs = self.ac.obs_oms(data['obs'].reshape(T * N, -1).float())       # pre_process_data
        act, adv = data['act'].reshape(T * N, -1), data['adv'].reshape(-1)
        target_v, logp_old = data['target_v'].reshape(-1), data['log_p'].reshape(-1)
        if self.use_standardized_advantages:
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        # ---- value net: mini-batch steps (iwpg.py:446-485)
        mbs = (T * N) // self.num_mini_batches
        for _ in range(self.train_v_iterations):
            perm = torch.randperm(T * N, device=obs.device)
            for start in range(0, mbs * self.num_mini_batches, mbs):
                idx = perm[start:start + mbs]
                self.vf_opt.zero_grad(set_to_none=True)
                loss_v = ((self.ac.v(obs[idx]).squeeze(-1) - target_v[idx]) ** 2).mean()
                loss_v.backward()
                self._avg_grads(self.ac.v)
                self.vf_opt.step()
        # ---- policy net: full-batch steps (iwpg.py:416-444)
        with torch.no_grad():
            mu_old, std_old, _ = self._log_prob(obs, act)
        kl, it = 0.0, 0
        for it in range(self.train_pi_iterations):
            self.pi_opt.zero_grad(set_to_none=True)
            loss_pi, mu, std = self._loss_pi(obs, act, adv, logp_old)
            loss_pi.backward()
            self._avg_grads(self.ac.pi)
            self.pi_opt.step()
            if self.use_kl_early_stopping:
                with torch.no_grad():
                    mu_n, std_n, _ = self._log_prob(obs, act)
                    k = (torch.log(std_n / std_old) + (std_old ** 2 + (mu_old - mu_n) ** 2) / (2 * std_n ** 2) - 0.5)
                    kl_t = k.sum(-1).mean().reshape(1)
                    if _world(self.dist) > 1:
                        self.dist.all_reduce(kl_t)
                        kl_t /= _world(self.dist)
                    kl = float(kl_t)
                if kl > self.target_kl:
                    break
        return {'loss_v': float(loss_v.detach()), 'loss_pi': float(loss_pi.detach()), 'pi_iters': it + 1, 'kl': kl}

    def learn_one_epoch(self):
        t0 = time.perf_counter()
        if self.use_exploration_noise_anneal:                # iwpg.py:272-273
            self.ac.set_log_std(1 - self.epoch / max(1, self.epochs))
        data = self.collector.collect()
        info = self.update(data)
        self.collector.update_running_statistics(data)       # after the update, on the raw data
        if self.sched is not None:
            self.sched.step()
        torch.cuda.synchronize(self.env.device)
        dt = time.perf_counter() - t0
        T, N = data['rew'].shape
        es = data['episode_stats']
        row = dict(epoch=self.epoch, EpRet=es.ret_mean, EpLen=es.len_mean, episodes=es.n,
                   FPS=T * N * _world(self.dist) / dt, **info)
        self.history.append(row)
        self.epoch += 1
        return row

    def learn(self, verbose=False):
        for _ in range(self.epochs):
            row = self.learn_one_epoch()
            if verbose:
                print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}, flush=True)
        return self.ac, self.env


def learn(env_id, **kwargs):
    """Counterpart of algs/ppo/ppo.py:50-63 `learn(env_id, **kwargs) -> (ac, env)`."""
    return PPO(env_id, **kwargs).learn()
