"""configs[4] through the fused collector kernel (pdx_collect) and through the two-kernel path, per policy precision.
    python tools/bench_collect.py [n_envs] [T] [rollouts]"""
import json
import sys

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from phoenix_drone_simulation_b200 import VecEnv
from phoenix_drone_simulation_b200.rollout import ActorCritic, RolloutCollector

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
R = int(sys.argv[3]) if len(sys.argv) > 3 else 5
for env_id in ('DroneHoverBulletEnv-v0', 'DroneHoverSimpleEnv-v0'):
    for kernel in ('tc', 'tc_tf32'):
        for fused in (True, False):
            torch.manual_seed(0)
            env = VecEnv(env_id, n, seed=2, keep_final_obs=True)
            ac = ActorCritic(env.obs_dim, policy_kernel=kernel)
            col = RolloutCollector(env, ac, T)
            col.use_fused_kernel = fused
            for _ in range(3):
                col.update_running_statistics(col.collect())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(R):
                data = col.collect()
                col.update_running_statistics(data)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(json.dumps({'env_id': env_id, 'policy_kernel': kernel, 'fused': fused and col.fused_used, 'envs': n, 'T': T,
                              'env_steps_per_s': R * T * n / (ms * 1e-3), 'ms_per_rollout': ms / R,
                              'episodes': data['episode_stats'].n}), flush=True)
            del env, ac, col
