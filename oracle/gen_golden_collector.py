"""TEST INFRASTRUCTURE ONLY -- golden vectors for the rollout-collector arithmetic, produced by the
UNMODIFIED reference (algs/core.py Buffer, utils/online_mean_std.py OnlineMeanStd,
utils/mpi_tools.py mpi_statistics_scalar), imported from /root/reference with the stand-ins of
oracle/shim/ for the packages this image lacks (mpi4py: a COMM_WORLD of one rank).

    python oracle/gen_golden_collector.py        # writes tests/golden_collector/*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get('PHOENIX_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'shim'))
sys.path.insert(0, REFERENCE)

from phoenix_drone_simulation.algs import core                              # noqa: E402
from phoenix_drone_simulation.utils.online_mean_std import OnlineMeanStd     # noqa: E402
from phoenix_drone_simulation.utils import mpi_tools                         # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden_collector')


class _AC:
    def __init__(self, ret_oms):
        self.ret_oms = ret_oms


def gae_case(seed, T, scaled):
    rng = np.random.default_rng(seed)
    ret_oms = OnlineMeanStd(shape=(1,))
    ret_oms.update(torch.as_tensor(rng.normal(-20, 15, 500).astype(np.float32)))
    buf = core.Buffer(_AC(ret_oms), (3,), (4,), T, 0.99, 0.95, 'gae', scaled, True, True)
    rew = rng.normal(-1.0, 2.0, T).astype(np.float32)
    rew[rng.random(T) < 0.1] -= 100.0                       # terminal penalties
    val = rng.normal(-5.0, 3.0, T).astype(np.float32)
    done = np.zeros(T, np.uint8)
    boot = np.zeros(T, np.float32)
    cuts = sorted(rng.choice(np.arange(2, T - 2), size=5, replace=False).tolist())
    for k, c in enumerate(cuts):
        done[c] = 1 if k % 2 == 0 else 2
        if done[c] == 2:
            boot[c] = np.float32(rng.normal(-5, 3))
    last_val = np.float32(rng.normal(-5, 3))
    for t in range(T):
        buf.store(obs=np.zeros(3, np.float32), act=np.zeros(4, np.float32), rew=rew[t], val=val[t], logp=0.0)
        if done[t] == 1:
            buf.finish_path(0.)                              # iwpg.py:379-380
        elif done[t] == 2:
            buf.finish_path(np.float32(boot[t]))             # iwpg.py:376-378
        elif t == T - 1:
            buf.finish_path(last_val)
    data = buf.get()
    return dict(rew=rew, val=val, done=done, boot_val=boot, last_val=last_val, scaled=np.bool_(scaled),
                ret_std=ret_oms.std.detach().numpy().copy(),
                adv=data['adv'].numpy(), target_v=data['target_v'].numpy(),
                disc_ret=data['discounted_ret'].numpy())


def oms_case(seed, dim, n_batches, rows):
    rng = np.random.default_rng(seed)
    oms = OnlineMeanStd(shape=(dim,))
    xs, means, stds, counts = [], [], [], []
    for b in range(n_batches):
        x = (rng.normal(0.3 * b, 1.0 + b, (rows, dim)) * np.linspace(0.1, 3, dim)).astype(np.float32)
        oms.update(torch.as_tensor(x) if dim > 1 else torch.as_tensor(x[:, 0]))
        xs.append(x)
        means.append(oms.mean.detach().numpy().copy())
        stds.append(oms.std.detach().numpy().copy())
        counts.append(oms.count.detach().numpy().copy())
    probe = rng.normal(0, 2, (7, dim)).astype(np.float32)
    fwd = oms(torch.as_tensor(probe)).numpy()
    fwd_clip = oms(torch.as_tensor(probe * 50), subtract_mean=False, clip=True).numpy()
    return dict(x=np.array(xs), mean=np.array(means), std=np.array(stds), count=np.array(counts),
                probe=probe, forward=fwd, forward_noclip_mean=fwd_clip)


def stats_case(seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(-60, 25, 37).astype(np.float32)
    mean, std, mn, mx = mpi_tools.mpi_statistics_scalar(x, with_min_and_max=True)
    return dict(x=x, mean=np.float64(mean), std=np.float64(std), min=np.float64(mn), max=np.float64(mx))


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, 'gae_scaled.npz'), **gae_case(1, 64, True))
    np.savez_compressed(os.path.join(OUT, 'gae_plain.npz'), **gae_case(2, 48, False))
    np.savez_compressed(os.path.join(OUT, 'oms_obs34.npz'), **oms_case(3, 34, 4, 96))
    np.savez_compressed(os.path.join(OUT, 'oms_ret1.npz'), **oms_case(4, 1, 3, 200))
    np.savez_compressed(os.path.join(OUT, 'stats.npz'), **stats_case(5))
    print('wrote', sorted(os.listdir(OUT)))


if __name__ == '__main__':
    main()
