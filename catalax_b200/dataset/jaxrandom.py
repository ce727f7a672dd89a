"""``jax.random.normal(jax.random.PRNGKey(seed), shape, float32)`` without jax, for
``Measurement.augment`` (reference: catalax/dataset/measurement.py:759-817 draws its noise from
jax's counter-based generator, so a drop-in has to produce the same numbers for the same seed).

jax's default generator is Threefry-2x32 with 20 rounds; since jax 0.5 (the reference pins
``jax>=0.7``) the 32 random bits of element ``i`` are the xor of the two output words of the block
with counter ``(0, i)`` under the key ``(seed >> 32, seed & 0xffffffff)``.  The uniform variate
keeps the top 23 bits as an f32 mantissa in ``[1, 2)``, and the normal one is
``sqrt(2) * erfinv(u)`` with ``u`` in ``(-1, 1)``.  Host-side numpy: this runs once per data set.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erfinv

_KS_PARITY = 0x1BD11BDA
_ROTATIONS = (13, 15, 26, 6, 17, 29, 16, 24)
_MASK = 0xFFFFFFFF


def _threefry_words(key0: int, key1: int, counters: np.ndarray):
    """Both 32-bit output words of the blocks with counters (0, counters[i])."""
    ks = (key0 & _MASK, key1 & _MASK, (key0 ^ key1 ^ _KS_PARITY) & _MASK)
    a = np.full(counters.shape, ks[0], dtype=np.uint64)
    b = (counters.astype(np.uint64) + ks[1]) & _MASK
    for rnd in range(20):
        rot = _ROTATIONS[(rnd // 4 % 2) * 4 + rnd % 4]
        a = (a + b) & _MASK
        b = (((b << rot) | (b >> (32 - rot))) & _MASK) ^ a
        if rnd % 4 == 3:
            inject = rnd // 4 + 1
            a = (a + ks[inject % 3]) & _MASK
            b = (b + ks[(inject + 1) % 3] + inject) & _MASK
    return a, b


def normal(seed: int, n: int, dtype=np.float32) -> np.ndarray:
    """n standard normal variates of ``PRNGKey(seed)`` in ``dtype``.

    f32 is pinned by known answers (tests/test_reference_pins.py).  f64 (data created under
    ``enable_x64``) follows the same recipe with 64 random bits per element -- first output word
    high, second low, 52 mantissa bits -- and has no published known answer to check against."""
    w0, w1 = _threefry_words(seed >> 32, seed, np.arange(n, dtype=np.uint32))
    if np.dtype(dtype) == np.float64:
        bits = (w0 << np.uint64(32)) | w1
        unit = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
        lo = np.nextafter(-1.0, 0.0)
        u = np.maximum(lo, unit * (1.0 - lo) + lo)
        return np.sqrt(2.0) * erfinv(u)
    bits = (w0 ^ w1).astype(np.uint32)
    unit = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = np.maximum(lo, unit * (np.float32(1.0) - lo) + lo)
    return (np.float32(np.sqrt(2.0)) * erfinv(u.astype(np.float64))).astype(np.float32)
