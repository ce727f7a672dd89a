"""Synthetic workloads of BASELINE.json's configs: model definitions and seeded
inputs shared by ``bench.py`` and the parity tests, so that the benchmark does not depend on
test code.  Models are ``(states, parameters, constants, odes, reactions)`` tuples of sympy
expressions; ``field_spec`` turns one into the code generator's input."""
from __future__ import annotations

import numpy as np
import sympy as sp

from .tools import codegen as cg

# config 2 parameter ranges: (K_m low, K_m high).  SURVEY §8d proposed LogU(0.01, 1): 3.6 % of
# such rows become stiff after substrate depletion (decay rate k_cat*E0/K_m up to 1000 per time
# unit) and exhaust max_steps=4096, which the reference's default throw=True turns into an
# exception; the headline range keeps every row solvable, the proposed one is reported beside it
# with throw=False semantics (inf fill).
KM_RANGE_SOLVABLE = (0.5, 50.0)
KM_RANGE_SURVEY = (0.01, 1.0)
# config 4 (NUTS): true parameters (K_m, k_cat).  With the README values (0.05, 0.1) the
# trajectories deplete and the DISCRETE tangents of the reference algorithm blow up;
# the benchmark posterior therefore sits at (40, 0.3).
NUTS_THETA_STAR = (40.0, 0.3)


def michaelis_menten():
    """README model in ODE form (reference README.md:99-116)."""
    S, E, P, Km, kc = sp.symbols("S E P K_m k_cat")
    r = kc * E * S / (Km + S)
    states, params, consts = ["E", "P", "S"], ["K_m", "k_cat"], []
    odes = {"E": sp.Integer(0), "P": r, "S": -r}
    return states, params, consts, odes, []


def mass_action_network(n_states=32):
    """ring of first-order steps + second-order couplings."""
    S = n_states
    states = [f"x{i:02d}" for i in range(S)]
    params = [f"k{j:02d}" for j in range(2 * S)]
    reactions = []
    for i in range(S):
        xi = sp.Symbol(states[i])
        reactions.append((f"r{i:02d}", [(1.0, states[i])], [(1.0, states[(i + 1) % S])],
                          sp.Symbol(params[i]) * xi))
    for i in range(S):
        xi, xj = sp.Symbol(states[i]), sp.Symbol(states[(i + 5) % S])
        reactions.append((f"r{S + i:02d}", [(1.0, states[i]), (1.0, states[(i + 5) % S])],
                          [(1.0, states[(i + 11) % S])], sp.Symbol(params[S + i]) * xi * xj))
    return states, params, [], {}, reactions


def field_spec(model):
    states, params, consts, odes, reactions = model
    return cg.FieldSpec(list(states), list(params), list(consts), dict(odes), list(reactions))


def mm_inputs(B, seed=0, dtype=np.float64, km_range=KM_RANGE_SOLVABLE):
    """Config 2 inputs (states ordered [E, P, S]): S0~U(50,300), E0~U(5,20), P0=0,
    k_cat~LogU(0.02,0.5), K_m~LogU(km_range)."""
    rng = np.random.default_rng(seed)
    S0 = rng.uniform(50, 300, B)
    E0 = rng.uniform(5, 20, B)
    y0 = np.stack([E0, np.zeros(B), S0], axis=1).astype(dtype)
    Km = np.exp(rng.uniform(np.log(km_range[0]), np.log(km_range[1]), B))
    kc = np.exp(rng.uniform(np.log(0.02), np.log(0.5), B))
    theta = np.stack([Km, kc], axis=1).astype(dtype)
    return y0, theta, np.zeros((B, 0), dtype=dtype)


def mass_action_inputs(B, seed=1, dtype=np.float64, n_states=32):
    """Config 3 inputs: x0~U(0.1,2), k~LogU(0.01,1) per row."""
    rng = np.random.default_rng(seed)
    y0 = rng.uniform(0.1, 2.0, (B, n_states)).astype(dtype)
    k = np.exp(rng.uniform(np.log(0.01), np.log(1.0), (B, 2 * n_states))).astype(dtype)
    return y0, k, np.zeros((B, 0), dtype=dtype)


def nuts_operands(chains, M, T, seed, clean_fn):
    """Config 4 operands: per-series y0 / time grid, per-chain theta ~ theta* LogN(0, 0.2) and
    sigma ~ LogN(log 0.5, 0.3); data = ``clean_fn(y0s, theta_star, ts)`` (noise-free trajectories
    at theta*, [M,T,3]) + N(0, 0.05 |y| + 1e-3) noise.  float64 arrays."""
    rng = np.random.default_rng(seed)
    y0s = np.stack([rng.uniform(5, 20, M), np.zeros(M), rng.uniform(50, 300, M)], axis=1)
    ts = np.tile(np.linspace(0, 100, T), (M, 1))
    theta_star = np.array(NUTS_THETA_STAR)
    clean = np.asarray(clean_fn(y0s, theta_star, ts), dtype=np.float64)
    data = clean + rng.normal(size=clean.shape) * (0.05 * np.abs(clean) + 1e-3)
    theta = theta_star * np.exp(rng.normal(0, 0.2, (chains, 2)))
    sigma = np.exp(rng.normal(np.log(0.5), 0.3, (chains, 3)))
    return y0s, theta, np.zeros((M, 0)), ts, data, sigma
