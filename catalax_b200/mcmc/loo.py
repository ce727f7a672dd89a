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
                               leave_out: str = "point", max_draws: Optional[int] = None,
                               config: Optional[SimulationConfig] = None, arrays=None):
    """loo.py:331-437 for mechanistic fits: per-observation, per-draw concentration-space
    log-likelihoods ready for ``arviz.loo``.  Returns ``(loglik [chains, draws, n_kept], posterior,
    meta)``.  (A surrogate-mode fit is scored by Euler-integrating its sampled rates in the
    reference, :203-303 -- host arithmetic on the K4 rate kernel's output, not built here.)"""
    from .predictive import subsample

    bm = results.bayesian_model
    if getattr(bm, "mode", "integrated") not in ("integrated", "mechanistic"):
        raise NotImplementedError("reconstruct_log_likelihood: surrogate-mode fits are not supported")
    posterior = subsample(results.mcmc.get_samples(group_by_chain=True), max_draws)
    ll, mask, info = mechanistic_loglik(results, dataset, posterior, sigma_source=sigma_source,
                                        yerrs=yerrs, config=config, arrays=arrays)
    loglik, meta = leave_out_groups(ll, mask, leave_out)
    names = [name for name, _ in bm.priors]
    posterior_np = {n: np.asarray(posterior[n]) for n in names if n in posterior}
    meta.update(sigma_source=sigma_source, mode="MECHANISTIC", integration="euler", info=info)
    return loglik, posterior_np, meta


def pointwise_elpd_is(loglik: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Plain importance-sampling LOO estimate per observation (the un-smoothed form of what
    arviz's PSIS computes from the same array): ``elpd_i = -log mean_s exp(-ll[s, i])``; NaN where
    masked.  For diagnostics and tests; use ``az.loo`` for the Pareto-smoothed statistic."""
    ll = loglik.reshape(-1, *mask.shape)
    m = (-ll).max(axis=0)
    elpd = -(m + np.log(np.mean(np.exp(-ll - m[None]), axis=0)))
    return np.where(mask, elpd, np.nan)
