#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (configs[2..4]) on ONE GPU -- companion of
bench.py (which measures configs[1], the headline).  Prints one JSON line per configuration;
the lines committed under profiles/ come from this script run under gpurun.

    python bench_configs.py [--scale 1.0] [--steps 20] [--warmup 3]
    torchrun --nproc-per-node N ... bench_configs.py --only "configs[4]"     # configs[4] on N GPUs (also: "configs[2]", "configs[3]")

configs[2]  DroneCircleSimpleEnv-v0, 524,288 envs per GPU (4 Mi over 8 GPUs), H = 2 and H = 8
configs[3]  DroneTakeOffSimpleEnv-v0, 1 Mi envs, ground effect on, obs noise, auto-reset (time limit
            and non-finite guard), bounded actions a = -0.1 + 0.1 N(0,1)
configs[4]  DroneHoverBulletEnv-v0 driving a PPO rollout (reference networks: pi 50-50 relu,
            v 64-64 tanh) with the device-side collector: policy forward, env.step, buffer
            writes, GAE and statistics, 131,072 envs x 64 steps per rollout
"""
import argparse
import json
import time

import torch

from phoenix_drone_simulation_b200 import VecEnv
from phoenix_drone_simulation_b200.rollout import ActorCritic, RolloutCollector


def open_loop(env_id, n, inner, steps, warmup, action_fn, **kw):
    """n environments PER GPU.  Under torchrun every rank steps its own shard (env_offset: the draws are those of one
    job of world x n environments); environments are independent, so the data path has no collective -- the episode
    statistics are summed once after the timed region; time = max over ranks."""
    dist, rank, world = _dist()
    dev = torch.device('cuda', torch.cuda.current_device())
    env = VecEnv(env_id, n, device=dev, seed=1, env_offset=rank * n, **kw)
    env.reset()
    dev, d = env.device, env.obs_dim
    g = torch.Generator(device=dev).manual_seed(0)
    acts = action_fn((2, inner, n, 4), dev, g)
    out = [{'obs': torch.empty((inner, n, d), device=dev), 'reward': torch.empty((inner, n), device=dev),
            'cost': torch.empty((inner, n), device=dev),
            'terminated': torch.empty((inner, n), dtype=torch.uint8, device=dev),
            'truncated': torch.empty((inner, n), dtype=torch.uint8, device=dev)} for _ in range(2)]
    for k in range(warmup):
        env.step_many(acts[k & 1], out[k & 1])
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        env.step_many(acts[k & 1], out[k & 1])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nonfinite = int((~torch.isfinite(out[0]['obs'])).sum()) + int((~torch.isfinite(out[0]['reward'])).sum())
    st = env.episode_stats()
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.cat([st[:4], torch.tensor([nonfinite], dtype=st.dtype, device=dev)])
        dist.all_reduce(tot)
        st = torch.cat([tot[:4], st[4:]])
        nonfinite = int(tot[4])
    s = st.cpu().tolist()
    bytes_per = env.rollout_bytes(inner) / inner
    n = n * world
    v = steps * inner * n / (ms * 1e-3)
    return {'env_id': env_id, 'envs': n, 'n_gpus': world, 'obs_dim': d, 'kwargs': {k: (v2 if not isinstance(v2, bool) else int(v2)) for k, v2 in kw.items()},
            'env_steps_per_s': v, 'ms_per_launch': ms / steps, 'env_steps_per_launch': inner * n,
            'algorithmic_bytes_per_env_step': bytes_per, 'algorithmic_GBps': v * bytes_per / 1e9,
            'nonfinite_words_in_last_segment': nonfinite, 'episodes_finished': int(s[0]), 'mean_episode_length': (s[3] / s[0]) if s[0] else None}


def uniform_actions(shape, dev, g):
    return torch.rand(shape, device=dev, generator=g) * 2 - 1


def takeoff_actions(shape, dev, g):
    return -0.1 + 0.1 * torch.randn(shape, device=dev, generator=g)


def _dist():
    """torch.distributed (NCCL) when launched under torchrun, else None.  One process per GPU."""
    import os
    if int(os.environ.get('WORLD_SIZE', '1')) == 1:
        return None, 0, 1
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG') and not os.environ.get('NCCL_DEBUG_FILE'):
            os.environ['NCCL_DEBUG_FILE'] = '/dev/stderr'     # stdout carries the JSON lines
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))
    return dist, dist.get_rank(), dist.get_world_size()


def ppo_rollout(n, T, rollouts, warmup, policy_kernel='tc', fused=True):
    """BASELINE configs[4]: n environments PER GPU; under torchrun every rank owns its shard (env_offset) and
    the running statistics / episode statistics are combined over NCCL inside the timed region; the
    time is the max over ranks."""
    dist, rank, world = _dist()
    torch.manual_seed(0)
    dev = torch.device('cuda', torch.cuda.current_device())
    env = VecEnv('DroneHoverBulletEnv-v0', n, device=dev, seed=2, keep_final_obs=True, env_offset=rank * n)
    ac = ActorCritic(env.obs_dim, device=env.device, policy_kernel=policy_kernel, dist=dist, seed=10000 * rank)
    col = RolloutCollector(env, ac, T, dist=dist)
    col.use_fused_kernel = fused
    for _ in range(warmup):
        data = col.collect()
        col.update_running_statistics(data)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rollouts):
        data = col.collect()
        col.update_running_statistics(data)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    wall = time.perf_counter() - t0
    es = data['episode_stats']
    n = n * world
    kern = {'tc': 'tcgen05 tensor-core kernel, split-TF32 (float32-level)', 'tc_tf32': 'tcgen05 tensor-core kernel, single TF32',
            'cuda': 'CUDA-core float32 kernel'}[policy_kernel]
    return {'env_id': 'DroneHoverBulletEnv-v0', 'envs': n, 'n_gpus': world, 'rollout_steps': T, 'policy_kernel': policy_kernel,
            'fused_collector_kernel': bool(col.fused_used),
            'what': (f'PPO rollout: pdx_collect, one persistent kernel per rollout (policy networks: {kern}; env.step with the state in '
                     'registers), GAE kernel, running statistics' if col.fused_used else
                     f'PPO rollout: policy step ({kern}) and env.step kernel alternating, GAE kernel, running-stat moments'),
            'env_steps_per_s': rollouts * T * n / (ms * 1e-3), 'ms_per_rollout': ms / rollouts, 'wall_s': wall,
            'episodes_in_last_rollout': es.n, 'ep_ret_mean': es.ret_mean, 'ep_len_mean': es.len_mean}


def ppo_training(n, T, epochs):
    """PPO end-to-end FPS (env-steps/s INCLUDING the SGD update) on
    DroneCircleBulletEnv-v0 with PWM control; the reference's logs: median 30,916 (unknown CPU host)."""
    from phoenix_drone_simulation_b200.ppo import PPO
    alg = PPO('DroneCircleBulletEnv-v0', num_envs=n, steps=T, epochs=epochs, seed=0)
    alg.learn()
    fps = sorted(r['FPS'] for r in alg.history[1:])
    return {'env_id': 'DroneCircleBulletEnv-v0', 'envs': n, 'rollout_steps': T, 'epochs': epochs,
            'what': 'PPO training: rollout (fused policy + env kernels, GAE) + torch update (80 full-batch policy '
                    'steps, 5 x 16 value mini-batch steps, iwpg.py defaults); env-steps/s including the update',
            'env_steps_per_s': fps[len(fps) // 2], 'published_reference_fps': 30916,
            'first_epoch': {k: alg.history[0][k] for k in ('EpRet', 'EpLen', 'episodes')},
            'last_epoch': {k: alg.history[-1][k] for k in ('EpRet', 'EpLen', 'episodes')}}


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--scale', type=float, default=1.0, help='scale the env counts (smoke runs)')
    p.add_argument('--steps', type=int, default=20)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--kernels', default='tc,tc_tf32,cuda', help='policy kernels of the configs[4] lines, in this order')
    p.add_argument('--only', default='', help='run only the lines whose config name contains this string')
    a = p.parse_args()
    sc = lambda n: max(1024, int(n * a.scale) // 128 * 128)
    import os
    emit = lambda ln: print(json.dumps(ln), flush=True) if int(os.environ.get('RANK', '0')) == 0 else None
    if a.only:
        lines = {'configs[2] H=2': lambda: emit(dict(config='configs[2] H=2', **open_loop(
                     'DroneCircleSimpleEnv-v0', sc(524288), 16, a.steps, a.warmup, uniform_actions))),
                 'configs[4]': lambda: [emit(dict(config=f'configs[4] policy_kernel={k}', **ppo_rollout(sc(131072), 64, max(2, a.steps // 5), 3, k)))
                                        for k in a.kernels.split(',')],
                 'configs[2] H=8': lambda: emit(dict(config='configs[2] H=8', **open_loop(
                     'DroneCircleSimpleEnv-v0', sc(524288), 8, a.steps, a.warmup, uniform_actions, observation_history_size=8))),
                 'configs[3]': lambda: emit(dict(config='configs[3]', **open_loop(
                     'DroneTakeOffSimpleEnv-v0', sc(1048576), 8, a.steps, a.warmup, takeoff_actions, use_ground_effect=True,
                     reset_on_nonfinite=True))),
                 'training': lambda: emit(dict(config='BASELINE.md PPO training FPS', **ppo_training(sc(16384), 64, 6)))}
        for name, fn in lines.items():
            if a.only in name:
                fn()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
        return
    emit(dict(config='configs[2] H=2', **open_loop('DroneCircleSimpleEnv-v0', sc(524288), 16, a.steps, a.warmup, uniform_actions)))
    emit(dict(config='configs[2] H=8', **open_loop('DroneCircleSimpleEnv-v0', sc(524288), 8, a.steps, a.warmup, uniform_actions,
                                                  observation_history_size=8)))
    emit(dict(config='configs[3]', **open_loop('DroneTakeOffSimpleEnv-v0', sc(1048576), 8, a.steps, a.warmup, takeoff_actions,
                                              use_ground_effect=True, reset_on_nonfinite=True)))
    emit(dict(config='configs[1] fixed policy', **open_loop('DroneHoverSimpleEnv-v0', sc(65536), 64, a.steps * 4, a.warmup,
                                                           lambda s, d, g: -0.1111 + 0.05 * torch.randn(s, device=d, generator=g))))
    emit(dict(config='configs[4] env only (open loop)', **open_loop('DroneHoverBulletEnv-v0', sc(131072), 32, a.steps, a.warmup,
                                                               lambda s, d, g: 0.1111 + 0.3 * torch.randn(s, device=d, generator=g))))
    for k in a.kernels.split(','):
        emit(dict(config=f'configs[4] policy_kernel={k}', **ppo_rollout(sc(131072), 64, max(2, a.steps // 5), 3, k)))
    emit(dict(config='BASELINE.md PPO training FPS', **ppo_training(sc(16384), 64, 6)))


if __name__ == '__main__':
    main()
