"""Developer probe: where the time of one PPO rollout step goes (policy kernel vs env kernel)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phoenix_drone_simulation_b200 import VecEnv
from phoenix_drone_simulation_b200.rollout import ActorCritic, RolloutCollector
n = 131072
env = VecEnv('DroneHoverBulletEnv-v0', n, seed=2, keep_final_obs=True)
ac = ActorCritic(env.obs_dim, device=env.device)
col = RolloutCollector(env, ac, 64)
for _ in range(2):
    d = col.collect(); col.update_running_statistics(d)
torch.cuda.synchronize()
def timeit(f, k=50):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k): f()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3, (t1 - t0) / k * 1e6
import ctypes as C
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
print('policy kernel us (gpu, host):', timeit(lambda: ac.step_prepared(col._prepared[3][0], stream)))
print('env step kernel us (gpu, host):', timeit(lambda: env.step_prepared(col._prepared[3][1], stream)))
print('collect ms (gpu, host us):', timeit(lambda: col.collect(), 5))
print('update stats ms:', timeit(lambda: col.update_running_statistics(d), 5))
envs = VecEnv('DroneHoverSimpleEnv-v0', n, seed=2)
a = torch.zeros((n, 4), device='cuda')
envs.reset()
print('hover simple step us:', timeit(lambda: envs.step(a)))
