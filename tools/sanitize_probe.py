"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every env family, fused
and single-step launches, ragged block, resets, wide rows, PID, collector kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phoenix_drone_simulation_b200 import VecEnv
from phoenix_drone_simulation_b200.rollout import ActorCritic, RolloutCollector
for env_id, kw in (('DroneHoverSimpleEnv-v0', {}), ('DroneCircleBulletEnv-v0', {'control_mode': 'AttitudeRate', 'aggregate_phy_steps': 4}),
                   ('DroneTakeOffSimpleEnv-v0', {'max_episode_steps': 5}), ('DroneCircleSimpleEnv-v0', {'observation_history_size': 8})):
    n, T = 333, 6
    env = VecEnv(env_id, n, seed=1, keep_final_obs=True, **kw)
    env.reset()
    g = torch.Generator(device='cuda').manual_seed(0)
    acts = (torch.rand((T, n, 4), device='cuda', generator=g) * 2 - 1).contiguous()
    for t in range(2):
        env.step(acts[t])
    out = {'obs': torch.zeros((T, n, env.obs_dim), device='cuda'), 'reward': torch.zeros((T, n), device='cuda'),
           'cost': torch.zeros((T, n), device='cuda'), 'terminated': torch.zeros((T, n), dtype=torch.uint8, device='cuda'),
           'truncated': torch.zeros((T, n), dtype=torch.uint8, device='cuda')}
    env.step_many(acts, out)
    torch.cuda.synchronize()
    print(env_id, 'ok', float(out['reward'].sum()))
env = VecEnv('DroneHoverBulletEnv-v0', 300, seed=2, keep_final_obs=True)
ac = ActorCritic(env.obs_dim, device='cuda')
col = RolloutCollector(env, ac, 5)
d = col.collect(); col.update_running_statistics(d)
torch.cuda.synchronize()
print('collector ok')
