"""Shared model definitions + seeded inputs for oracle-vs-CUDA parity tests."""
from __future__ import annotations

import numpy as np
import sympy as sp

from catalax_b200.tools import codegen as cg  # noqa: F401
from catalax_b200.workloads import (field_spec, mass_action_inputs, mass_action_network,  # noqa: F401
                                    michaelis_menten, mm_inputs)


def menten_fixture():
    """tests/conftest.py:8-19 of the reference: 1 state, constant e."""
    s1, e, kcat, km = sp.symbols("s1 e kcat k_m")
    return ["s1"], ["k_m", "kcat"], ["e"], {"s1": kcat * e * s1 / (km + s1)}, []


def nonautonomous():
    """4-state field that exercises t, exp, pow, sqrt, *_init, constants."""
    a, b, c, d, k1, k2, k3, q, t = sp.symbols("a b c d k1 k2 k3 q t")
    a_init = sp.Symbol("a_init")
    odes = {
        "a": -k1 * a * sp.exp(-k3 * t) + q * sp.sqrt(b + 1),
        "b": k1 * a - k2 * b ** 2 / (1 + b),
        "c": k2 * b * b / (1 + b) - k3 * c * sp.sin(t),
        "d": k3 * (a_init - d) + sp.tanh(c),
    }
    return ["a", "b", "c", "d"], ["k1", "k2", "k3"], ["q"], odes, []


def mm_closed_form(t, s0, V, K):
    """S(t) of dS/dt = -V S/(K+S): u = S/K solves u + ln u = ln(s0/K) + (s0 - V t)/K
    (Lambert W written so that it cannot overflow for small K)."""
    L = np.log(s0 / K) + (s0 - V * np.asarray(t, dtype=np.float64)) / K
    u = np.where(L > 1.0, L - np.log(np.maximum(L, 1.0)), np.exp(np.minimum(L, 1.0)))
    for _ in range(60):
        u = np.maximum(u, 1e-300)
        u = u - (u + np.log(u) - L) / (1.0 + 1.0 / u)
    return K * np.maximum(u, 0.0)
