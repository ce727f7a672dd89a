"""numpy restatement of ``jax.random.normal(jax.random.PRNGKey(seed), (n,), float32)``.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Needed to reproduce the data of the
reference's ``examples/HMC.ipynb``: ``Dataset.augment(seed=0, sigma=5e-2, multiplicative=True)``
(``catalax/dataset/measurement.py:793-817``) multiplies every series by
``1 + 0.05 * jax.random.normal(PRNGKey(0), x.shape, x.dtype)`` -- the SAME key for every
measurement and state (``dataset.py:1238-1246`` passes ``seed + i`` with ``i`` the augmentation
index), so one vector of ``T`` normals defines the whole augmented data set.

jax is an un-vendored dependency (``jax>=0.7.0,<0.8.0``, ``pyproject.toml``); restated from its
published algorithm:

* ``PRNGKey(seed)`` = the uint32 pair ``(seed >> 32, seed & 0xffffffff)``;
* random bits: Threefry-2x32, 20 rounds (Salmon et al., SC'11; rotations 13,15,26,6 / 17,29,16,24,
  key-schedule parity 0x1BD11BDA).  With ``jax_threefry_partitionable=True`` (the default from
  jax 0.5) element ``i`` of a 32-bit draw is ``x0 ^ x1`` of ``threefry(key, (hi(i), lo(i)))``;
  the legacy layout hashes ``iota(n)`` (zero-padded to even length) split in two halves and
  concatenates the two output words;
* ``uniform``: ``(bits >> 9) | 0x3f800000`` viewed as f32, minus 1, scaled to ``[lo, 1)`` with
  ``lo = nextafter(-1, 0)``; ``normal = sqrt(2) * erfinv(u)``.

Known answers (``tests/test_reference_pins.py``): Random123's ``threefry2x32_20`` vector for a
zero key / counter, and the values jax's documentation prints for ``random.uniform`` /
``random.normal`` of ``key(0)`` and ``key(42)`` in both layouts.  erfinv is evaluated in f64 and
rounded (XLA uses an f32 polynomial): differences are a few f32 ulps of the noise, 1e-7 of the
data.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erfinv

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32-20 of counter words (x0, x1) under key (k0, k1); uint32 arrays."""
    x0 = np.array(x0, dtype=np.uint32, ndmin=1)
    x1 = np.array(x1, dtype=np.uint32, ndmin=1)
    k0, k1 = np.uint32(k0), np.uint32(k1)
    ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
    with np.errstate(over="ignore"):
        x0 = (x0 + ks[0]).astype(np.uint32)
        x1 = (x1 + ks[1]).astype(np.uint32)
        for blk in range(5):
            for r in _ROT[blk % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r) ^ x0
            x0 = (x0 + ks[(blk + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(blk + 2) % 3] + np.uint32(blk + 1)).astype(np.uint32)
    return x0, x1


def random_bits32(seed: int, n: int, partitionable: bool = True) -> np.ndarray:
    k0, k1 = (seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF
    if partitionable:
        a, b = threefry2x32(k0, k1, np.zeros(n, np.uint32), np.arange(n, dtype=np.uint32))
        return a ^ b
    count = np.arange(n, dtype=np.uint32)
    if n % 2:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    half = count.size // 2
    a, b = threefry2x32(k0, k1, count[:half], count[half:])
    return np.concatenate([a, b])[:n]


def uniform(seed: int, n: int, partitionable: bool = True) -> np.ndarray:
    """jax.random.uniform(PRNGKey(seed), (n,), float32) in [0, 1)."""
    bits = random_bits32(seed, n, partitionable)
    return ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)


def normal(seed: int, n: int, partitionable: bool = True) -> np.ndarray:
    """jax.random.normal(PRNGKey(seed), (n,), float32)."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = np.maximum(lo, uniform(seed, n, partitionable) * (np.float32(1.0) - lo) + lo)
    return (np.float32(np.sqrt(2.0)) * erfinv(u.astype(np.float64))).astype(np.float32)


def augment_multiplicative(data: np.ndarray, sigma: float, seed: int,
                           partitionable: bool = True) -> np.ndarray:
    """``Measurement._jitter_data(x, sigma, seed, multiplicative=True)`` for every series of
    ``data [M, T]`` (f32): ``x * (normal * sigma + 1)``."""
    data = np.asarray(data, dtype=np.float32)
    z = normal(seed, data.shape[-1], partitionable)
    return data * (z * np.float32(sigma) + np.float32(1.0))
