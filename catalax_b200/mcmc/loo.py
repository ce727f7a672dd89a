"""Mechanistic point-wise log-likelihood of posterior draws -- the expensive input of PSIS-LOO.

Mirror of ``catalax/mcmc/loo.py:73-200`` (``_build_eval_config``, ``_mechanistic_loglik``) and of
``reconstruct_log_likelihood`` (:331-437: point / curve leave-out bookkeeping) for mechanistic fits: the
reference re-integrates the ODE once per posterior draw (``jax.vmap(per_draw)(thetas, sigma_y)``)
and scores every observation under the model's own likelihood.  Here that is ONE fused kernel
launch over the two-level (draw x series) batch whose dense-output pass writes the log-density
of each observation (``ctx_pointwise_loglik``).  The PSIS smoothing itself is arviz's
(``az.loo`` on the returned array) and is not part of this engine."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from ..model.simconfig import SimulationConfig
from .mcmc import extract_dataset_components
from .predictive import EnsembleIntegrator, sigma_y_draws, stack_thetas


def build_eval_config(results) -> SimulationConfig:
    """loo.py:73-89: the fit's solver and dt0, throw=False (failed solves give non-finite
    states that the mask drops)."""
    bm = results.bayesian_model
    return SimulationConfig(t1=1, t0=0, dt0=bm.config.dt0, solver=bm.config.solver, throw=False)


def mechanistic_loglik(results, dataset, posterior: Dict[str, np.ndarray], *,
                       sigma_source: str = "reuse", yerrs=None,
                       config: Optional[SimulationConfig] = None, arrays=None):
    """loo.py:131-200.  Returns ``(loglik [n_chain, n_draw, M, T, n_obs], mask [M,T,n_obs], info)``.

    ``sigma_source="reuse"``: the sampled concentration-space ``sigma_y`` is the noise scale;
    ``"yerrs"``: score against the supplied measurement error (scalar or per-observable).
    ``arrays``: optional ``(data, times, y0s, constants)`` instead of a Dataset (array seam)."""
    bm = results.bayesian_model
    config = config or build_eval_config(results)
    data, times, y0s, constants = arrays if arrays is not None else \
        extract_dataset_components(dataset, results.model)
    mask = np.isfinite(data)                                     # mcmc.py:566
    data_safe = np.where(mask, data, 0.0)
    n_obs = data.shape[-1]
    thetas, (n_chain, n_draw) = stack_thetas(posterior, [name for name, _ in bm.priors])
    if sigma_source == "yerrs":
        if yerrs is None:
            raise ValueError("sigma_source='yerrs' needs yerrs")
        if np.ndim(yerrs) > 1:      # loo.py:306-328 also broadcasts a full [M,T,n_obs] error array
            raise NotImplementedError("per-point yerrs: the fused kernel takes one scale per observable")
        sig = np.broadcast_to(np.asarray(yerrs, dtype=np.float64).reshape(-1)[-n_obs:] if np.ndim(yerrs)
                              else np.full(n_obs, float(yerrs)), (n_chain * n_draw, n_obs)).copy()
    elif sigma_source == "reuse":
        sig = sigma_y_draws(posterior, n_chain, n_draw, n_obs)
    else:
        raise ValueError(f"unknown sigma_source {sigma_source!r}")
    integ = EnsembleIntegrator(results.model, config)
    if n_obs != integ.field.S:
        raise ValueError("observables must be all modeled states (mcmc.py:393-395, 772-775)")
    ll = integ.pointwise_log_likelihood(y0s, thetas, sig, constants, times, data_safe,
                                        likelihood=bm.likelihood)
    ll = ll.cpu().numpy().reshape(n_chain, n_draw, *data.shape)
    info = {"data": np.asarray(data), "times": np.asarray(times),
            "species": list(results.model.get_observable_state_order())}
    return ll, mask, info


def _log_prob(likelihood: str, x, loc, scale):
    """numpyro's Normal / SoftLaplace log-density (the two likelihoods of the reference)."""
    z = (x - loc) / scale
    if likelihood == "normal":
        return -0.5 * z * z - np.log(scale) - 0.5 * np.log(2 * np.pi)
    if likelihood == "softlaplace":
        return np.log(2 / np.pi) - np.log(scale) - np.logaddexp(z, -z)
    raise ValueError(f"unknown likelihood {likelihood!r}")


def surrogate_loglik(results, dataset, posterior: Dict[str, np.ndarray], *, sigma_source: str = "reuse",
                     yerrs=None, integration: str = "euler", arrays=None, draw_chunk: int = 256):
    """loo.py:203-303 for a surrogate-mode (rate-matching) fit: the mechanistic rates of every
    posterior draw at the measured states (ONE ``ctx_rates`` launch per chunk of draws: draws x
    points with per-point parameters), forward-Euler integrated along the measurement times
    (``predictive.euler_integrate_rates``) and scored against the measured concentrations.
    ``arrays``: optional ``(full_data [M,T,S], times [M,T], y0s (unused), constants [M,C])``.
    Returns ``(loglik [n_chain, n_draw, M, T, n_obs], mask, info)``."""
    import torch

    from .. import engine
    from ..config import default_dtype
    from ..tools import codegen as cg
    from .predictive import euler_integrate_rates

    bm, model = results.bayesian_model, results.model
    if arrays is not None:
        full_data, full_times, _, constants = arrays
    else:
        full_data, full_times, _, constants = extract_dataset_components(dataset, model)
    full_data, full_times = np.asarray(full_data, dtype=np.float64), np.asarray(full_times, dtype=np.float64)
    M, T, S = full_data.shape
    mask = np.isfinite(full_data)
    data_safe = np.where(mask, full_data, 0.0)
    y0_full = full_data[:, 0, :]
    dt = np.diff(full_times, axis=-1)
    thetas, (n_chain, n_draw) = stack_thetas(posterior, [name for name, _ in bm.priors])
    if sigma_source == "reuse":
        scale = sigma_y_draws(posterior, n_chain, n_draw, S)[:, None, None, :]
    elif sigma_source == "yerrs":
        if yerrs is None:       # loo.py:316-325: the stored error of a surrogate fit is rate-shaped
            raise ValueError("`yerrs` was not provided: pass concentration-space `yerrs` explicitly")
        scale = np.broadcast_to(np.asarray(yerrs, dtype=np.float64), (M, T, S))[None]
    else:
        raise ValueError(f"unknown sigma_source {sigma_source!r}")
    dtype = default_dtype()
    compiled = engine.CompiledModel(
        cg.generate_field(model._replace_assignments().sim_input.field_spec()), np.dtype(dtype).name)
    dev = torch.device("cuda", torch.cuda.current_device())
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=dtype), device=dev)
    N = M * T
    pts, t_flat = tt(full_data.reshape(N, S)), tt(full_times.reshape(N))
    c_flat = tt(np.repeat(np.asarray(constants, dtype=np.float64).reshape(M, -1), T, axis=0))
    inits = tt(np.repeat(y0_full, T, axis=0))              # "{state}_init" = first measured state
    D = thetas.shape[0]
    rates = np.empty((D, M, T, S))
    for lo in range(0, D, draw_chunk):
        th = tt(thetas[lo:lo + draw_chunk])
        d = th.shape[0]
        r = compiled.rates(t_flat.repeat(d), pts.repeat(d, 1), th.repeat_interleave(N, dim=0),
                           c_flat.repeat(d, 1), y0=inits.repeat(d, 1))
        rates[lo:lo + d] = r.reshape(d, M, T, S).cpu().numpy()
    yhat, _ = euler_integrate_rates(rates, y0_full, dt, np.zeros((M, T, S)), integration=integration,
                                    data_safe=data_safe)
    with np.errstate(all="ignore"):
        ll = _log_prob(bm.likelihood, data_safe[None], yhat, scale)
    ll = ll.reshape(n_chain, n_draw, M, T, S)
    info = {"data": full_data, "times": full_times, "species": list(model.get_state_order())}
    return ll, mask, info


def leave_out_groups(ll: np.ndarray, mask: np.ndarray, leave_out: str = "point"):
    """loo.py:395-426: turn ``ll [n_chain, n_draw, M, T, n_obs]`` into the held-out units arviz
    scores.  ``"point"``: one column per OBSERVED (series, time, observable) entry; ``"curve"``: one
    column per series that has data, masked points contributing 0.  Returns ``(loglik [n_chain,
    n_draw, n_kept], meta)``."""
    n_chain, n_draw = ll.shape[:2]
    event_shape = ll.shape[2:]
    if leave_out == "point":
        keep = mask.reshape(-1)
        kept_idx = np.flatnonzero(keep)
        loglik = ll.reshape(n_chain, n_draw, -1)[:, :, kept_idx]
        meta = {"leave_out": "point", "kept_idx": kept_idx, "full_shape": event_shape,
                "n_total": int(keep.size), "n_dropped": int(keep.size - kept_idx.size)}
    elif leave_out == "curve":
        n_meas = event_shape[0]
        per_curve = np.where(mask[None, None], ll, 0.0).reshape(n_chain, n_draw, n_meas, -1).sum(axis=-1)
        kept_idx = np.flatnonzero(mask.reshape(n_meas, -1).any(axis=-1))
        loglik = per_curve[:, :, kept_idx]
        meta = {"leave_out": "curve", "kept_idx": kept_idx, "full_shape": (n_meas,),
                "n_total": int(n_meas), "n_dropped": int(n_meas - kept_idx.size)}
    else:
        raise ValueError(f"Unknown leave_out={leave_out!r}; use 'point' or 'curve'.")
    return loglik, meta


def reconstruct_log_likelihood(results, dataset, *, yerrs=None, sigma_source: str = "reuse",
                               leave_out: str = "point", integration: str = "euler",
                               max_draws: Optional[int] = None,
                               config: Optional[SimulationConfig] = None, arrays=None):
    """loo.py:331-437: per-observation, per-draw concentration-space log-likelihoods ready for
    ``arviz.loo``, for mechanistic fits (one fused solve + likelihood launch) and surrogate-mode
    fits (rates at the measured states, Euler-integrated: ``integration`` = ``"euler"`` |
    ``"euler_onestep"``).  Returns ``(loglik [chains, draws, n_kept], posterior, meta)``."""
    from .predictive import subsample

    bm = results.bayesian_model
    posterior = subsample(results.mcmc.get_samples(group_by_chain=True), max_draws)
    surrogate = getattr(bm, "mode", "integrated") == "surrogate"
    if surrogate:
        ll, mask, info = surrogate_loglik(results, dataset, posterior, sigma_source=sigma_source,
                                          yerrs=yerrs, integration=integration, arrays=arrays)
    else:
        ll, mask, info = mechanistic_loglik(results, dataset, posterior, sigma_source=sigma_source,
                                            yerrs=yerrs, config=config, arrays=arrays)
    loglik, meta = leave_out_groups(ll, mask, leave_out)
    names = [name for name, _ in bm.priors]
    posterior_np = {n: np.asarray(posterior[n]) for n in names if n in posterior}
    meta.update(sigma_source=sigma_source, mode="SURROGATE" if surrogate else "MECHANISTIC",
                integration=integration, info=info)
    return loglik, posterior_np, meta


def pointwise_elpd_is(loglik: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Plain importance-sampling LOO estimate per observation (the un-smoothed form of what
    arviz's PSIS computes from the same array): ``elpd_i = -log mean_s exp(-ll[s, i])``; NaN where
    masked.  For diagnostics and tests; use ``az.loo`` for the Pareto-smoothed statistic."""
    ll = loglik.reshape(-1, *mask.shape)
    m = (-ll).max(axis=0)
    elpd = -(m + np.log(np.mean(np.exp(-ll - m[None]), axis=0)))
    return np.where(mask, elpd, np.nan)
