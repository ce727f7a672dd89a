"""CPU restatement of the NUTS log-likelihood and its gradient.  PARITY ORACLE (test infra).

Follows ``catalax/mcmc/mcmc.py:405-491`` (``BayesianModel.__call__``):

    states = sim_func(y0s, theta, constants, times)      # :448, in_axes=(0, None, 0, 0) (:860)
    loc = states[..., observables]                        # :461, observables = all modeled states
    sigma_total = sigma_y ** 2                            # :483
    with mask(mask): sample("y", likelihood(loc, sqrt(sigma_total)), obs=data)   # :486-491

and numpyro 0.19's distributions (un-vendored dependency, ``pyproject.toml:17``):
``Normal.log_prob = -((x-loc)/s)^2/2 - log s - log sqrt(2 pi)`` and
``SoftLaplace.log_prob = log(2/pi) - log s - logaddexp(z, -z)``, ``z = (x-loc)/s``.
The gradient is what ``jax.value_and_grad`` of that log-density gives: the chain rule through
the discrete tangents of the solve (``diffrax_oracle.solve`` with ``n_err``), ``d/dsigma``
through ``s = sqrt(sigma^2)``.  PARITY: numpyro cannot be run here and the reference asserts no
log-density value in its tests; the DENSITY is pinned through the posterior
summaries two real NUTS runs of the reference printed (``examples/HMC.ipynb``,
``examples/InhomogeneousData.ipynb``; ``tests/test_reference_pins.py``), the GRADIENT against
finite differences of the frozen-step oracle and closed-form sensitivities (unpinned against
``jax.grad`` itself).
"""
from __future__ import annotations

import numpy as np

from . import diffrax_oracle as do, symbolic as sy


def logpdf_terms(x, loc, s, likelihood):
    z = (x - loc) / s
    if likelihood == "normal":
        lp = -0.5 * z * z - np.log(s) - 0.5 * np.log(2 * np.pi)
        dl_dloc = z / s
        dl_ds = (z * z - 1.0) / s
    elif likelihood == "softlaplace":
        lp = np.log(2 / np.pi) - np.log(s) - np.logaddexp(z, -z)
        dl_dloc = np.tanh(z) / s
        dl_ds = (z * np.tanh(z) - 1.0) / s
    else:
        raise ValueError(likelihood)
    return lp, dl_dloc, dl_ds


def loglik_grad(model, y0s, theta, consts, ts, data, sigma, *, mask=None,
                likelihood="softlaplace", sens_y0=False, dtype=np.float64, **solve_kw):
    """model = (states, parameters, constants, odes, reactions); theta [CH,P]; sigma [CH,S].

    Returns out [CH, 1+P+S], grad_y0 [CH, M, S] (zeros unless sens_y0), status [CH, M].
    """
    states, params, consts_n, odes, reactions = model
    S, P = len(states), len(params)
    rhs, N, D, init_aug = sy.make_rhs(states, params, consts_n, odes, reactions, sens_theta=True,
                                      sens_y0=sens_y0)
    y0s = np.asarray(y0s, dtype=dtype); ts = np.asarray(ts, dtype=dtype)
    data = np.asarray(data, dtype=dtype); consts = np.asarray(consts, dtype=dtype)
    M, T = ts.shape
    if mask is None:
        mask = np.isfinite(data)
    CH = theta.shape[0]
    out = np.zeros((CH, 1 + P + S), dtype=np.float64)
    gy0 = np.zeros((CH, M, S), dtype=np.float64)
    status = np.zeros((CH, M), dtype=np.int32)
    for ch in range(CH):
        r = do.solve(rhs, y0s, np.asarray(theta[ch], dtype=dtype), consts, ts, dtype=dtype,
                     n_err=S, init_aug=init_aug(M, dtype), **solve_kw)
        status[ch] = r.status
        ys = r.ys[..., :S].astype(np.float64)
        tang = r.ys[..., S:].astype(np.float64).reshape(M, T, D, S)   # [M,T,dir,state]
        s = np.abs(np.asarray(sigma[ch], dtype=np.float64))
        with np.errstate(all="ignore"):
            lp, dl_dloc, dl_ds = logpdf_terms(np.where(mask, data, 0.0).astype(np.float64), ys,
                                              s[None, None, :], likelihood)
            out[ch, 0] = np.where(mask, lp, 0.0).sum()
            w = np.where(mask, dl_dloc, 0.0)
            for p in range(P):
                out[ch, 1 + p] = (w * tang[:, :, p, :]).sum()
            out[ch, 1 + P:] = (np.where(mask, dl_ds, 0.0) * np.sign(sigma[ch])[None, None, :]).sum(axis=(0, 1))
            if sens_y0:
                for i in range(S):
                    gy0[ch, :, i] = (w * tang[:, :, P + i, :]).sum(axis=(1, 2))
    return out, gy0, status


def loglik_grad_c(model, y0s, theta, consts, ts, data, sigma, *, mask=None,
                  likelihood="softlaplace", dtype=np.float64, n_threads=None, oracle=None,
                  **solve_kw):
    """``loglik_grad`` with the solves done by the compiled C / OpenMP oracle (one task per
    (chain, series) trajectory with its parameter tangents): the CPU baseline of bench.py's NUTS
    leg and the checker for chain counts the numpy form is too slow for.  Same restated
    algorithm (``c_oracle.py``); out [CH, 1+P+S] float64, status [CH, M]."""
    from .c_oracle import COracle

    states, params, consts_n, odes, reactions = model
    S, P = len(states), len(params)
    orc = oracle or COracle(states, params, consts_n, odes, reactions, sens_theta=True, dtype=dtype)
    y0s = np.asarray(y0s, dtype=dtype); ts = np.asarray(ts, dtype=dtype)
    data = np.asarray(data, dtype=np.float64); consts = np.asarray(consts, dtype=dtype)
    theta = np.asarray(theta, dtype=dtype)
    M, T = ts.shape
    CH = theta.shape[0]
    if mask is None:
        mask = np.isfinite(data)
    r = orc.solve(np.tile(y0s, (CH, 1)), np.repeat(theta, M, axis=0), np.tile(consts, (CH, 1)),
                  np.tile(ts, (CH, 1)), n_threads=n_threads, **solve_kw)
    ys = r.ys.reshape(CH, M, T, -1)
    loc = ys[..., :S].astype(np.float64)
    tang = ys[..., S:].astype(np.float64).reshape(CH, M, T, P, S)
    sg = np.asarray(sigma, dtype=np.float64)
    s = np.abs(sg)[:, None, None, :]
    with np.errstate(all="ignore"):
        lp, dl_dloc, dl_ds = logpdf_terms(np.where(mask, data, 0.0)[None], loc, s, likelihood)
        out = np.zeros((CH, 1 + P + S))
        out[:, 0] = np.where(mask[None], lp, 0.0).sum(axis=(1, 2, 3))
        w = np.where(mask[None], dl_dloc, 0.0)
        out[:, 1:1 + P] = np.einsum("cmts,cmtps->cp", w, tang)
        out[:, 1 + P:] = np.where(mask[None], dl_ds, 0.0).sum(axis=(1, 2)) * np.sign(sg)
    return out, r.status.reshape(CH, M), r


def surrogate_loglik_grad(model, t, y, theta, consts, inits, data, jac2, sigma, *, n_time,
                          mask=None, extra2=None, likelihood="softlaplace"):
    """Surrogate (rate-matching) mode of the same numpyro model.  PARITY ORACLE (test infra).

    Follows ``catalax/mcmc/mcmc.py``:
      * :819-847  ``sim_func = vmap(rate_fun, in_axes=(0, None, 0, 0, 0))``: the mechanistic
        right-hand side evaluated at the measured states ``y [N,S]`` (``N = M * n_time`` points,
        grouped per measurement), constants repeated per time point, ``{state}_init`` symbols
        from each measurement's first state (``inits [M,S]``); ``data`` = surrogate rates;
      * :469-480  ``sigma_total[p,j] = sum_k J[p,j,k]^2 sigma_y[k]^2 (+ extra) + 1e-8``;
      * :486-491  masked ``likelihood(loc, sqrt(sigma_total))``.
    Gradient = chain rule through ``d rate / d theta`` (sympy derivative, numpy evaluation) and
    through ``sqrt(sigma_total)``.  Returns out [CH, 1+P+S] (float64).
    """
    states, params, consts_n, odes, reactions = model
    S, P = len(states), len(params)
    rhs, _N, D, _ = sy.make_rhs(states, params, consts_n, odes, reactions, sens_theta=True)
    t = np.asarray(t, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    N = y.shape[0]
    M = N // n_time
    rows = np.repeat(np.arange(M), n_time)
    c_pts = np.asarray(consts, dtype=np.float64).reshape(M, -1)[rows]
    i_pts = np.asarray(inits, dtype=np.float64)[rows]
    data = np.asarray(data, dtype=np.float64)
    jac2 = np.asarray(jac2, dtype=np.float64)
    if mask is None:
        mask = np.isfinite(data)
    ex = np.zeros((N, S)) if extra2 is None else np.asarray(extra2, dtype=np.float64)
    CH = theta.shape[0]
    out = np.zeros((CH, 1 + P + S))
    Y = np.concatenate([y, np.zeros((N, D * S))], axis=1)      # zero tangents: ds = d f / d theta
    for ch in range(CH):
        th = np.broadcast_to(np.asarray(theta[ch], dtype=np.float64), (N, P))
        r = rhs(t, Y, th, c_pts, i_pts)
        loc = r[:, :S]
        dloc = r[:, S:].reshape(N, D, S)                         # [N, dir, state]
        sg = np.asarray(sigma[ch], dtype=np.float64)
        s2 = np.einsum("pjk,k->pj", jac2, sg ** 2) + ex + 1e-8
        sd = np.sqrt(s2)
        with np.errstate(all="ignore"):
            lp, dl_dloc, dl_ds = logpdf_terms(np.where(mask, data, 0.0), loc, sd, likelihood)
        out[ch, 0] = np.where(mask, lp, 0.0).sum()
        w = np.where(mask, dl_dloc, 0.0)
        for p in range(P):
            out[ch, 1 + p] = (w * dloc[:, p, :]).sum()
        ws = np.where(mask, dl_ds / sd, 0.0)                     # d lp / d s2 * 2 ... via sd
        out[ch, 1 + P:] = np.einsum("pj,pjk->k", ws, jac2) * sg
    return out
