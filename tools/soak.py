"""Developer soak run: thousands of fused steps per env family at full size; checks finiteness,
flag / statistics consistency and the time-limit bookkeeping."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phoenix_drone_simulation_b200 import VecEnv
N, T, R = 65536, 64, 40
for env_id, kw, amp in (('DroneHoverSimpleEnv-v0', {}, 0.05), ('DroneCircleBulletEnv-v0', {}, 0.1),
                        ('DroneTakeOffSimpleEnv-v0', {'reset_on_nonfinite': True}, 0.1),
                        ('DroneHoverBulletEnv-v0', {'control_mode': 'Attitude', 'aggregate_phy_steps': 4}, 0.2),
                        ('DroneCircleSimpleEnv-v0', {'observation_history_size': 8}, 1.0)):
    env = VecEnv(env_id, N, seed=1, **kw)
    env.reset()
    g = torch.Generator(device='cuda').manual_seed(0)
    out = {'obs': torch.zeros((T, N, env.obs_dim), device='cuda'), 'reward': torch.zeros((T, N), device='cuda'),
           'cost': torch.zeros((T, N), device='cuda'), 'terminated': torch.zeros((T, N), dtype=torch.uint8, device='cuda'),
           'truncated': torch.zeros((T, N), dtype=torch.uint8, device='cuda'), 'episode_length': torch.zeros((T, N), dtype=torch.int32, device='cuda')}
    n_fin, n_trunc, bad, max_len = 0, 0, 0, 0
    for r in range(R):
        base = env.cfg.hover_action if 'TakeOff' not in env_id else -0.1
        acts = (base + amp * torch.randn((T, N, 4), device='cuda', generator=g)).contiguous()
        env.step_many(acts, out)
        fin = (out['terminated'] | out['truncated']) > 0
        n_fin += int(fin.sum()); n_trunc += int(out['truncated'].sum())
        bad += int((~torch.isfinite(out['obs'])).sum())
        max_len = max(max_len, int(out['episode_length'].max()))
        assert (out['episode_length'][fin] >= 1).all() and (out['episode_length'][~fin] == 0).all()
    s = env.episode_stats().cpu().tolist()
    assert int(s[0]) == n_fin, (s[0], n_fin)
    assert max_len <= 500 and s[7] <= 500
    print(f'{env_id} {kw}: {R*T} steps x {N} envs, episodes {n_fin}, truncated {n_trunc}, max len {max_len}, '
          f'non-finite obs words {bad}, mean return {s[1]/max(1,s[0]):.2f}')
