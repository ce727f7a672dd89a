"""Posterior re-integration: one solve per posterior draw x measurement series.

Mirror of ``catalax/mcmc/predictive.py:109-155`` (``_integrate_ensemble``:
``jax.vmap(per_draw)(thetas)`` over ``sim_func(y0s, theta, constants, times)`` ->
``[n_draws, n_meas, n_steps, n_obs]``), ``:158-175`` (``_aleatoric_variance``), ``:212-249``
(``compute_ensemble``) and of the mechanistic point-wise log-likelihood that LOO uses
(``catalax/mcmc/loo.py:131-200``, see ``loo.py`` here).  All are the same CUDA kernel as
``Model.simulate`` with a two-level batch (``ctx_solve_cfg.inner``): parameters are indexed by
draw, initial values / constants / times by series, so nothing is tiled or copied on the host;
the log-likelihood form never materialises the ``[D,M,T,S]`` trajectories (ctx_pointwise_loglik).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from ..config import default_dtype
from ..solvers import Tsit5, solver_name
from ..tools import codegen as cg


def ensemble_times(times: np.ndarray, n_steps: int) -> np.ndarray:
    """Per-series dense grid t0 + (t1 - t0) * linspace(0, 1, n_steps) (predictive.py:140-143)."""
    times = np.asarray(times)
    t0, t1 = times.min(axis=-1), times.max(axis=-1)
    frac = np.linspace(0.0, 1.0, n_steps)
    return (t0[:, None] + (t1 - t0)[:, None] * frac[None, :]).astype(times.dtype)


def stack_thetas(posterior: Dict[str, np.ndarray], param_names) -> Tuple[np.ndarray, Tuple[int, int]]:
    """Parameter draws ``(n_chain, n_draw, ...)`` -> ``(n_chain*n_draw, n_params)`` in prior order
    (loo.py:92-99)."""
    cols = [np.asarray(posterior[name]) for name in param_names]
    n_chain, n_draw = cols[0].shape[:2]
    flat = [c.reshape(n_chain * n_draw, *c.shape[2:]) for c in cols]
    return np.stack(flat, axis=-1), (n_chain, n_draw)


def sigma_y_draws(posterior, n_chain: int, n_draw: int, n_obs: int) -> np.ndarray:
    """Flattened per-observable ``sigma_y`` draws ``(n_chain*n_draw, n_obs)``; a shared scalar is
    broadcast, a missing site gives ones (loo.py:102-117)."""
    if "sigma_y" in posterior:
        flat = np.asarray(posterior["sigma_y"]).reshape(n_chain * n_draw, -1)
        return np.broadcast_to(flat, (n_chain * n_draw, n_obs)).copy()
    return np.ones((n_chain * n_draw, n_obs))


def subsample(posterior, max_draws: Optional[int]):
    """Evenly subsample the draw axis (axis 1) of a grouped-by-chain sample dict (loo.py:120-128)."""
    if max_draws is None:
        return posterior
    n_draw = next(iter(posterior.values())).shape[1]
    if n_draw <= max_draws:
        return posterior
    idx = np.linspace(0, n_draw - 1, max_draws).astype(int)
    return {k: np.asarray(v)[:, idx] for k, v in posterior.items()}


class EnsembleIntegrator:
    """Integrates a mechanistic model for every (draw, series) pair on the GPU."""

    def __init__(self, model, config=None):
        from .. import engine
        from ..model.simconfig import SimulationConfig

        self.model = model
        # predictive.py:126-128: Tsit5, throw=False, default tolerances
        self.config = config or SimulationConfig(t1=1, t0=0, throw=False)
        self.dtype = default_dtype()
        spec = model._replace_assignments().sim_input.field_spec()
        self.field = cg.generate_field(spec)
        self.compiled = engine.CompiledModel(self.field, np.dtype(self.dtype).name)
        self.observables = list(range(self.field.S))

    def _tensors(self, *arrays, device=None):
        import torch

        dev = device or torch.device("cuda", torch.cuda.current_device())
        tt = lambda a: (a if isinstance(a, torch.Tensor) else
                        torch.as_tensor(np.asarray(a, dtype=self.dtype))).to(dev)
        return [tt(a) for a in arrays]

    def integrate(self, y0s, thetas, constants, times, device=None):
        """y0s [M,S], thetas [D,P], constants [M,C], times [M,T] -> torch CUDA [D,M,T,S]."""
        y0s, thetas, constants, times = self._tensors(y0s, thetas, constants, times, device=device)
        M, T = times.shape
        D = thetas.shape[0]
        cfg = self.config
        out = self.compiled.solve(y0s, thetas, constants.reshape(M, self.field.C), times,
                                  solver=solver_name(cfg.solver), in_axes=(0, 0, 0, 0),
                                  rtol=cfg.rtol, atol=cfg.atol, dt0=cfg.dt0,
                                  max_steps=cfg.max_steps, inner=M)
        self.last = out
        return out.ys.reshape(D, M, T, self.field.S)

    def pointwise_log_likelihood(self, y0s, thetas, sigmas, constants, times, data,
                                 likelihood="normal", device=None):
        """log p(data[m,t,o] | draw d) for every draw and observation -> torch [D,M,T,S]: the input
        of LOO's PSIS (loo.py:131-200).  One fused launch (ctx_pointwise_loglik): the dense-output
        pass of the solve evaluates the likelihood of the observation at each requested time, the
        ``[D,M,T,S]`` trajectories are never written.  Missing observations (NaN in ``data``) are
        scored on 0 inside the kernel (``data_safe``, loo.py:158) and come back as NaN; failed solves
        give -inf / NaN like the reference's inf states."""
        import torch

        y0s, thetas, sigmas, constants, times, data = self._tensors(
            y0s, thetas, sigmas, constants, times, data, device=device)
        M, T = times.shape
        D = thetas.shape[0]
        cfg = self.config
        missing = ~torch.isfinite(data)                       # mcmc.py:566: the mask is isfinite(data)
        data_safe = torch.where(missing, torch.zeros_like(data), data)
        out = self.compiled.pointwise_loglik(
            y0s, thetas, constants.reshape(M, self.field.C), times, data_safe, sigmas,
            likelihood=likelihood, solver=solver_name(cfg.solver), rtol=cfg.rtol, atol=cfg.atol,
            dt0=cfg.dt0, max_steps=cfg.max_steps, inner=M)
        self.last = out
        ll = out.ys.reshape(D, M, T, self.field.S)
        if bool(missing.any()):
            ll = torch.where(missing[None].expand_as(ll), torch.full_like(ll, float("nan")), ll)
        return ll


def integrate_ensemble(results, dataset, posterior, *, n_steps: int, arrays=None):
    """``_integrate_ensemble`` (predictive.py:109-155): Tsit5 solve of ``results.model`` once per
    posterior draw on the dense per-series grid.  Returns ``(times [M,n_steps],
    yhat_ens [n_draws, M, n_steps, n_obs])`` (numpy).  ``arrays``: optional ``(data, times, y0s,
    constants)`` instead of a Dataset."""
    from ..model.simconfig import SimulationConfig
    from .mcmc import extract_dataset_components

    bm = results.bayesian_model
    config = SimulationConfig(t1=1, t0=0, dt0=bm.config.dt0, solver=Tsit5, throw=False)
    _, meas_times, y0s, constants = arrays if arrays is not None else \
        extract_dataset_components(dataset, results.model)
    times = ensemble_times(meas_times, n_steps)
    thetas, _ = stack_thetas(posterior, [name for name, _ in bm.priors])
    integ = EnsembleIntegrator(results.model, config)
    ys = integ.integrate(y0s, thetas, constants, times)
    obs = results.model.get_observable_state_order(as_indices=True)
    return times, ys[..., obs].cpu().numpy()


def euler_integrate_rates(rates, y0_obs, dt, rate_var, *, integration: str = "euler", data_safe=None):
    """predictive.py:54-106: forward-Euler integration of observable rates into concentrations
    (shared by the surrogate-mode LOO and the predictive bands).  ``rates`` / ``rate_var``
    ``[..., M, T, n_obs]`` (any leading draw axes), ``y0_obs [M, n_obs]``, ``dt [M, T-1]``;
    ``"euler"`` integrates from the measured initial state, ``"euler_onestep"`` predicts every point
    from the previous MEASURED state (``data_safe``).  Returns ``(yhat, var_y)``."""
    rates, rate_var = np.asarray(rates), np.asarray(rate_var)
    lead = rates.shape[:-3]
    y0b = np.broadcast_to(np.asarray(y0_obs)[:, None, :], lead + (rates.shape[-3], 1, rates.shape[-1]))
    step = rates[..., :-1, :] * np.asarray(dt)[..., None]
    var_inc = (np.asarray(dt)[..., None] ** 2) * rate_var[..., :-1, :]
    zeros = np.zeros(lead + (rates.shape[-3], 1, rates.shape[-1]))
    if integration == "euler_onestep":
        yhat = np.concatenate([y0b, np.asarray(data_safe)[..., :-1, :] + step], axis=-2)
        var_y = np.concatenate([zeros, np.broadcast_to(var_inc, step.shape)], axis=-2)
    else:
        yhat = np.concatenate([y0b, y0b + np.cumsum(step, axis=-2)], axis=-2)
        var_y = np.concatenate([zeros, np.cumsum(np.broadcast_to(var_inc, step.shape), axis=-2)], axis=-2)
    return yhat, var_y


def aleatoric_variance(posterior, yhat_shape) -> np.ndarray:
    """Per-draw concentration-space aleatoric variance broadcast to ``yhat_shape``
    (predictive.py:158-175): ``sigma_y**2``; a shared scalar is broadcast across observables."""
    n_draws = yhat_shape[0]
    sigma_y = np.asarray(posterior["sigma_y"]).reshape(n_draws, -1)
    return np.broadcast_to((sigma_y ** 2)[:, None, None, :], yhat_shape).copy()


def compute_ensemble(results, dataset, *, n_steps: int = 100, max_draws: Optional[int] = None,
                     yerrs=None):
    """``compute_ensemble`` (predictive.py:212-249): push the posterior through the model.
    Returns ``(times [M,n_steps], yhat_ens, var_ens)``, the ensembles ``[n_draws, M, n_steps, n_obs]``."""
    posterior = subsample(results.get_samples(group_by_chain=True), max_draws)
    times, yhat_ens = integrate_ensemble(results, dataset, posterior, n_steps=n_steps)
    if yerrs is not None:
        var_ens = np.broadcast_to(np.asarray(yerrs, dtype=np.float64) ** 2, yhat_ens.shape).copy()
    else:
        var_ens = aleatoric_variance(posterior, yhat_ens.shape)
    return times, yhat_ens, var_ens


#: HDI band identifier -> percentile of the predictive distribution (predictive.py:43-51)
PERCENTILE = {"lower": 2.5, "upper": 97.5, "lower_50": 25.0, "upper_50": 75.0}


def band_values(yhat_ens, var_ens, *, hdi: Optional[str] = None, include_noise: bool = True) -> np.ndarray:
    """predictive.py:252-293 without the Dataset wrapping: the posterior-median trajectory
    (``hdi=None``) or the percentile band ``[M, n_steps, n_obs]`` of the ensemble.  With
    ``include_noise`` the aleatoric component is sampled per draw and pooled in -- the noise is
    ``jax.random.normal(PRNGKey(0), shape)`` in the reference, reproduced here by the Threefry
    restatement of ``dataset/jaxrandom.py`` so that the same ensemble gives the same band."""
    yhat_ens = np.asarray(yhat_ens)
    if hdi is None:
        return np.nanmedian(yhat_ens, axis=0)
    if hdi not in PERCENTILE:
        raise KeyError(hdi)
    pooled = yhat_ens
    if include_noise:
        from ..dataset.jaxrandom import normal

        noise = normal(0, yhat_ens.size, np.float64 if default_dtype() == np.float64 else np.float32)
        pooled = yhat_ens + np.sqrt(np.asarray(var_ens)) * noise.reshape(yhat_ens.shape)
    return np.nanpercentile(pooled, PERCENTILE[hdi], axis=0)


def ensemble_to_dataset(reference, results, times, values):
    """predictive.py:177-209: wrap ``values [M, n_time, n_obs]`` on ``times [M, n_time]`` into a
    Dataset whose measurement ids and initial conditions are copied from ``reference``."""
    from ..dataset import Dataset, Measurement

    obs = list(results.model.get_observable_state_order())
    times, values = np.asarray(times), np.asarray(values)
    out = Dataset(states=reference.states, id=reference.id, name=reference.name)
    for i, ref in enumerate(reference.measurements):
        out.add_measurement(Measurement(id=ref.id, initial_conditions=dict(ref.initial_conditions),
                                        time=times[i],
                                        data={s: values[i, :, k] for k, s in enumerate(obs)}))
    return out


def band_from_ensemble(results, reference, times, yhat_ens, var_ens, *, hdi: Optional[str] = None,
                       include_noise: bool = True):
    """predictive.py:252-293."""
    return ensemble_to_dataset(reference, results, times,
                               band_values(yhat_ens, var_ens, hdi=hdi, include_noise=include_noise))


def posterior_predictive(results, dataset, *, hdi: Optional[str] = None, n_steps: int = 100,
                         max_draws: Optional[int] = None, include_noise: bool = True, yerrs=None):
    """predictive.py:296-331: ``compute_ensemble`` + ``band_from_ensemble`` in one call."""
    times, yhat_ens, var_ens = compute_ensemble(results, dataset, n_steps=n_steps, max_draws=max_draws,
                                                yerrs=yerrs)
    return band_from_ensemble(results, dataset, times, yhat_ens, var_ens, hdi=hdi,
                              include_noise=include_noise)
