import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phoenix_drone_simulation_b200 import VecEnv
N, T = 131072, 16
env = VecEnv('DroneHoverBulletEnv-v0', N, seed=1)
env.reset()
g = torch.Generator(device='cuda').manual_seed(0)
acts = (0.1111 + 0.3 * torch.randn((T, N, 4), device='cuda', generator=g)).contiguous()
out = {'obs': torch.zeros((T, N, env.obs_dim), device='cuda'), 'reward': torch.zeros((T, N), device='cuda'),
       'cost': torch.zeros((T, N), device='cuda'), 'terminated': torch.zeros((T, N), dtype=torch.uint8, device='cuda'),
       'truncated': torch.zeros((T, N), dtype=torch.uint8, device='cuda')}
for k in range(4):
    env.step_many(acts, out)
torch.cuda.synchronize()
