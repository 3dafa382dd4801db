"""Where a rollout's time goes: pdx_collect launch alone, the rest of collect(), update_running_statistics."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phoenix_drone_simulation_b200 import VecEnv
from phoenix_drone_simulation_b200.rollout import ActorCritic, RolloutCollector
from torch.profiler import profile, ProfilerActivity
env_id, kernel, n, T = sys.argv[1], sys.argv[2], int(sys.argv[3]), 64
torch.manual_seed(0)
env = VecEnv(env_id, n, seed=2, keep_final_obs=True)
ac = ActorCritic(env.obs_dim, policy_kernel=kernel)
col = RolloutCollector(env, ac, T)
for _ in range(3):
    col.update_running_statistics(col.collect())
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        col.update_running_statistics(col.collect())
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]
tot = sum(e.device_time_total for e in prof.key_averages())
print(env_id, kernel, 'total device us per rollout', round(tot / 3, 1))
for e in rows:
    print(f'{e.device_time_total / 3:9.1f} us  x{e.count // 3:3d}  {e.key[:90]}')
