"""Phase timeline of k_policy_tc (python -m phoenix_drone_simulation_b200.build --timing builds the -DPDX_TC_TIMING library): clock64 stamps of warp 2 of CTA 0
for its first eight tiles.  python tools/policy_tc_timing.py [tc|tc_tf32]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phoenix_drone_simulation_b200 import lib
lib.LIB_PATH = os.path.join(os.path.dirname(lib.LIB_PATH), 'libphoenix_b200_timing.so')   # built with -DPDX_TC_TIMING
from phoenix_drone_simulation_b200.rollout import ActorCritic

kernel = sys.argv[1] if len(sys.argv) > 1 else 'tc'
n, d = 1048576, 34
ac = ActorCritic(d, 4, device='cuda', policy_kernel=kernel, seed=1)
obs = torch.randn((n, d), device='cuda')
act = torch.zeros((n, 4), device='cuda'); val = torch.zeros(n, device='cuda'); logp = torch.zeros(n, device='cuda')
for _ in range(3):
    ac.step_into(obs, act, val, logp)
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
L = lib.load()
assert L.pdx_policy_tc_timing(buf) == 0
names = ['loop top', 'L1 done', 'a2 handed', 'x built', 'L2 done', 'epi2 done', 'synced', 'out done']
for t in range(1, 6):
    row = [buf[t * 16 + k] for k in range(8)]
    base = row[0]
    print(f'tile {t}: ' + '  '.join(f'{names[k]}={row[k] - base}' for k in range(1, 8)) + f'  | next top={buf[(t + 1) * 16] - base}')
