"""TEST INFRASTRUCTURE ONLY -- numpy restatement of Philox4x32-10 (Salmon, Moraes, Dror, Shaw:
"Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 reference implementation),
used to check the CUDA generator (csrc/pdx_math.cuh) bit for bit through pdx_dump_draws.
Known-answer vectors: Random123 `kat_vectors` (philox4x32 10 ...), see tests/test_philox.py.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32 array-like, key: (..., 2).  Returns (..., 4) uint32."""
    c = np.array(ctr, dtype=np.uint64) & MASK
    k = np.array(key, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


# draw-site ids of csrc/pdx_math.cuh (DrawSite)
SITE_SUBSTEP, SITE_FINAL_OBS, SITE_RESET, SITE_DR = 0, 64, 80, 88
SITE_RESET_OBS1, SITE_RESET_OBS2, SITE_INIT = 96, 104, 112


def engine_raw(seed, env_index, counter, site):
    """The 4 raw words the engine draws at `site` for global env `env_index` at `counter`
    (make_rng / Rng::raw in csrc/pdx_kernels.cuh, csrc/pdx_math.cuh)."""
    env_index, counter, seed = int(env_index), int(counter), int(seed)
    env_lo, env_hi = env_index & 0xFFFFFFFF, (env_index >> 32) & 0xFFFFFFFF
    env_hi ^= ((counter >> 32) << 8) & 0xFFFFFFFF
    ctr = [env_lo, counter & 0xFFFFFFFF, env_hi, site]
    key = [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]
    return philox4x32_10(ctr, key)


def uniforms_f64(raw):
    return raw.astype(np.float64) * (1.0 / 4294967296.0)


def normals_f64(raw):
    u = (raw.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    r0, r1 = np.sqrt(-2.0 * np.log(u[..., 0])), np.sqrt(-2.0 * np.log(u[..., 2]))
    a0, a1 = 2 * np.pi * u[..., 1], 2 * np.pi * u[..., 3]
    return np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], axis=-1)
