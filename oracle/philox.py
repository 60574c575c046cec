"""Philox4x32-10 counter-based RNG in numpy.  TEST INFRASTRUCTURE (oracle).

Restates the published Philox4x32-10 algorithm (Salmon et al., "Parallel random
numbers: as easy as 1, 2, 3", SC'11; Random123 ``philox.h``).  The reference
(IC3Net) draws from numpy MT19937 / torch CPU generators, which cannot be
reproduced on a GPU (SURVEY.md section 7 "RNG parity"); the CUDA kernels use this
counter-based generator instead and the oracle consumes the very same stream, so
"same seed" parity is exact.  Stream layout (shared with
``ic3net_b200/csrc/ic3_rng.cuh``):

    key     = (seed_lo, seed_hi)
    counter = (env_global_id, tick, stream, index)

    stream 1  PP reset     tick = episode counter, index = attempt block
    stream 2  TJ spawn     tick = env step counter, index = arrival group;
                           words: [spawn test, dead-slot choice, path choice, -]
    stream 3  action       tick = env step counter, index = agent;
                           words: one per action head (<= 4 heads)

Every draw is reduced to 24 bits (``w >> 8``) so that the uniform
``u = u24 * 2**-24`` is exact in fp32 and fp64 and all comparisons / index
selections are integer operations on both sides.
"""
import numpy as np

STREAM_PP_RESET = 1
STREAM_TJ_SPAWN = 2
STREAM_ACTION = 3

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32(counter, key):
    """counter: [...,4] uint32-like, key: (k0, k1) ints.  Returns [...,4] uint32."""
    c = np.asarray(counter, dtype=np.uint64) & _MASK
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def split_seed(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def draw_u24(seed, env_id, tick, stream, index):
    """Four 24-bit integers for one counter value."""
    w = philox4x32(np.array([env_id, tick, stream, index], dtype=np.uint64), split_seed(seed))
    return (w >> np.uint32(8)).astype(np.int64)


def u24_to_float(u24):
    return np.asarray(u24, dtype=np.float64) * (2.0 ** -24)


def float_to_u24(u):
    """Inverse of u24_to_float for tape values (exact multiples of 2**-24)."""
    v = np.asarray(u, dtype=np.float64) * (2.0 ** 24)
    r = np.floor(v).astype(np.int64)
    return r
