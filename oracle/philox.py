"""Counter-based random draws shared by the oracle and the CUDA kernels (TEST INFRASTRUCTURE).

The reference draws its randomness inside TensorFlow (``tfd.TruncatedNormal.sample`` at
``careless/models/merging/surrogate_posteriors.py:50-53`` -> a rejection sampler, and
``scale_dist.sample`` at ``careless/models/merging/variational.py:157`` -> Philox normals).
Neither stream can be reproduced outside TF, so parity is defined on *injected* draws:
uniforms ``u_f`` in (0,1) for the structure factors (inverse-CDF truncated normal) and
standard normals ``eps_s`` for the scales.  When nothing is injected the CUDA kernels
generate the draws in-kernel with Philox4x32-10 keyed by (seed; index, sample, step,
stream); this module restates that generator in numpy so the oracle can follow the very
same stream and parity also holds in the production (non-injected) mode.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)

STREAM_REFL = 0  # uniforms for the truncated-normal surrogate
STREAM_OBS = 1   # normals for the per-observation scale sample


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u01(x):
    """uint32 -> float in (0,1) on the 23-bit grid ((x>>9)+0.5)/2^23, exact in float32 (24 significant bits)."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) + 0.5) / 8388608.0


def _key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def refl_uniforms(seed, step, n_samples, refl_index):
    """u_f[s, r] for global reflection indices ``refl_index`` (shape (R,)). Returns float64 (S, R)."""
    k0, k1 = _key(seed)
    idx = np.asarray(refl_index, dtype=np.uint32)[None, :]
    s = np.arange(n_samples, dtype=np.uint32)[:, None]
    x0, _, _, _ = philox4x32_10(idx, s, np.uint32(step & 0xFFFFFFFF), np.uint32(STREAM_REFL), k0, k1)
    return u01(x0)


def obs_normals(seed, step, n_samples, obs_index):
    """eps_s[s, i] for global observation indices ``obs_index``: Box-Muller on two Philox words."""
    k0, k1 = _key(seed)
    idx = np.asarray(obs_index, dtype=np.uint32)[None, :]
    s = np.arange(n_samples, dtype=np.uint32)[:, None]
    x0, x1, _, _ = philox4x32_10(idx, s, np.uint32(step & 0xFFFFFFFF), np.uint32(STREAM_OBS), k0, k1)
    u1 = u01(x0)
    u2 = u01(x1)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
