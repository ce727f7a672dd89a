"""The NUTS log-density + gradient entry point.

Mirror of ``catalax/mcmc/mcmc.py``: ``MCMCConfig`` (:38-79), ``BayesianModel`` (:332-491),
``_prepare_mcmc_data`` (:539-592), ``_extract_dataset_components`` (:759-790).  The reference
hands a numpyro model to NUTS, which calls ``value_and_grad`` of its log-joint once per leapfrog
step; the sampler itself (numpyro) is not part of this engine.  What is provided is that
log-joint and its gradient, for MANY chains at once:

    log p(theta, sigma_y | data) = sum_p log prior_p(theta_p) + log prior(sigma_y)
                                   + sum_{mask} log L(data; loc = ys[..., obs], scale = |sigma_y|)

The likelihood term and its gradient come from the fused CUDA kernel (ctx_loglik_grad); priors are
scalar torch ops.  Multi-GPU: shard="chains" needs no collective; shard="series" all-reduces
the [chains, 1+P+n_obs] block (catalax_b200.parallel).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

from ..config import default_dtype
from ..model.simconfig import SimulationConfig
from ..solvers import Tsit5, solver_name
from . import priors as _priors
from .error import DEFAULT_ERROR_MODEL, ErrorModel


@dataclass
class MCMCConfig:
    """Fields and defaults of the reference's MCMCConfig (mcmc.py:57-71)."""

    num_warmup: int
    num_samples: int
    likelihood: Any = "softlaplace"      # reference: dist.SoftLaplace (class); "normal" also used
    dense_mass: bool = True
    thinning: int = 1
    max_tree_depth: int = 10
    dt0: float = 0.1
    chain_method: str = "sequential"
    num_chains: int = 1
    seed: int = 420
    verbose: int = 1
    max_steps: int = 64 ** 4
    solver: Any = Tsit5
    target_accept_prob: float = 0.8
    error_model: Optional[ErrorModel] = None

    def to_simulation_config(self) -> SimulationConfig:
        # NB the reference forwards max_steps as t1 (mcmc.py:73-79): the solves really run
        # with SimulationConfig's own defaults max_steps=4096, rtol=atol=1e-5.
        return SimulationConfig(t1=self.max_steps, t0=0, dt0=self.dt0, solver=self.solver)


def likelihood_name(lik) -> str:
    name = lik if isinstance(lik, str) else getattr(lik, "__name__", type(lik).__name__)
    name = name.lower()
    if name not in ("normal", "softlaplace"):
        raise NotImplementedError(f"likelihood {lik!r}: Normal and SoftLaplace are supported")
    return name


def extract_dataset_components(dataset, model):
    """(data [M,T,n_obs], times [M,T], y0s [M,S], constants [M,C]) -- mcmc.py:759-790."""
    data, times, _ = dataset.to_jax_arrays(model.get_observable_state_order(), inits_to_array=True)
    constants = dataset.to_y0_matrix(state_order=model.get_constants_order())
    order = model.get_state_order()
    y0s = np.stack([np.array([m.initial_conditions[s] for s in order], dtype=default_dtype())
                    for m in dataset.measurements])
    return data, times, y0s, constants


def surrogate_rate_jacobian(surrogate, states_flat, times_flat):
    """``J[p, i, k] = d rate_i / d y_k`` of a surrogate at every measured point
    (``_surrogate_rate_jacobian``, mcmc.py:865-894; computed once, the surrogate is fixed).

    Uses ``surrogate.rate_jacobian(t, y)`` when the surrogate provides one, otherwise central
    differences of its public ``rates`` API (the reference takes ``jax.jacfwd`` of the same
    function; the result feeds a first-order delta method, mcmc.py:469-480)."""
    y = np.asarray(states_flat, dtype=np.float64)
    t = np.asarray(times_flat, dtype=np.float64)
    if hasattr(surrogate, "rate_jacobian"):
        return np.asarray(surrogate.rate_jacobian(t, y), dtype=np.float64)
    N, S = y.shape
    J = np.empty((N, S, S))
    probe = np.asarray(surrogate.rates(t, y))
    eps = np.finfo(probe.dtype if probe.dtype.kind == "f" else np.float64).eps ** (1.0 / 3.0)
    for k in range(S):
        h = eps * np.maximum(1.0, np.abs(y[:, k]))
        yp, ym = y.copy(), y.copy()
        yp[:, k] += h
        ym[:, k] -= h
        J[:, :, k] = (np.asarray(surrogate.rates(t, yp), dtype=np.float64)
                      - np.asarray(surrogate.rates(t, ym), dtype=np.float64)) / (2.0 * h)[:, None]
    return J


def per_species_scale_hint(yerrs, n_obs: int) -> np.ndarray:
    """``yerrs`` -> default noise scale per observable, shape (n_obs,) (mcmc.py:316-329): the mean
    over every leading axis when the last axis lines up with the observables (a scalar gives the
    same value for every species, a full ``[M, T, n_obs]`` array its per-species means), else the
    global mean broadcast across species."""
    arr = np.asarray(yerrs, dtype=np.float64)
    if arr.ndim >= 1 and arr.shape[-1] == n_obs:
        return arr.reshape(-1, n_obs).mean(axis=0)
    return np.full(n_obs, arr.mean())


_per_species_scale_hint = per_species_scale_hint        # the reference's (private) name


class BayesianModel:
    """Batched log-density of the posterior the reference samples with NUTS (mcmc.py:332-491).

    ``surrogate`` switches to the reference's Modes.SURROGATE (mcmc.py:819-847, 469-480): no
    solve; the mechanistic rates at the measured states are matched to the surrogate's rates,
    with the concentration noise ``sigma_y`` pushed forward through the surrogate's Jacobian."""

    def __init__(self, model, dataset=None, *, yerrs=1.0, likelihood="softlaplace",
                 config: Optional[SimulationConfig] = None, error_model: Optional[ErrorModel] = None,
                 estimate_initials: bool = False, shard: str = "chains", arrays=None,
                 device=None, surrogate=None, surrogate_arrays=None):
        """arrays: optional (data, times, y0s, constants) instead of a Dataset (array seam).
        surrogate: object with ``predict_rates(dataset)`` / ``rates(t, y)`` (catalax/surrogate.py);
        surrogate_arrays: optional (rates [N,S], jac [N,S,S], extra2 [N,S] | None) instead."""
        from .. import engine, parallel
        from ..tools import codegen as cg

        self.model = model
        self.config = config or MCMCConfig(0, 0).to_simulation_config()
        self.likelihood = likelihood_name(likelihood)
        self.error_model = error_model or DEFAULT_ERROR_MODEL
        self.shard = shard
        if shard not in ("chains", "series", "none"):
            raise ValueError("shard must be 'chains', 'series' or 'none'")
        self.dtype = default_dtype()
        data, times, y0s, constants = arrays if arrays is not None else \
            extract_dataset_components(dataset, model)
        if data.shape[-1] != len(model.get_state_order()):
            raise ValueError("observables must be all modeled states (mcmc.py:393-395, 772-775)")
        self.priors = []
        for name in model.get_parameter_order():
            pr = model.parameters[name].prior
            if isinstance(pr, dict):
                pr = _priors.from_dict(pr)
            if pr is None:
                raise ValueError(f"Parameter '{name}' has no prior")
            self.priors.append((name, pr))
        rank, world = parallel.world()
        self.rank, self.world = rank, world
        M = data.shape[0]
        lo, hi = parallel.shard_bounds(M, rank, world) if shard == "series" else (0, M)
        self._series = slice(lo, hi)
        self._arrays = tuple(np.asarray(a, dtype=self.dtype)[lo:hi] for a in (data, times, y0s, constants))
        self.mask = np.isfinite(self._arrays[0])                       # mcmc.py:566
        self.yerrs = per_species_scale_hint(yerrs, data.shape[-1])
        self.mode = "surrogate" if (surrogate is not None or surrogate_arrays is not None) else "integrated"
        self._surr = None
        if self.mode == "surrogate":
            if estimate_initials:
                raise NotImplementedError("surrogate mode evaluates rates at the measured states; "
                                          "initial conditions are not sampled")
            T_, S_ = data.shape[1], data.shape[2]
            if surrogate_arrays is not None:
                rates, jac, extra2 = surrogate_arrays
            else:
                if dataset is None:
                    raise ValueError("a surrogate needs the Dataset (predict_rates) or surrogate_arrays")
                rates = np.asarray(surrogate.predict_rates(dataset), dtype=np.float64)
                jac = surrogate_rate_jacobian(surrogate, np.asarray(data).reshape(-1, S_),
                                              np.asarray(times).ravel())
                extra2 = None
                if getattr(surrogate, "has_uncertainty", False):      # mcmc.py:225-230, 472-475
                    extra2 = np.asarray(surrogate.rate_uncertainty(dataset=dataset)).reshape(-1, S_) ** 2 \
                        + np.asarray(surrogate.rate_sigma(dataset=dataset)).reshape(1, S_) ** 2
            rates = np.asarray(rates, dtype=self.dtype).reshape(M, T_, S_)[lo:hi].reshape(-1, S_)
            jac2 = (np.asarray(jac, dtype=np.float64) ** 2).astype(self.dtype) \
                .reshape(M, T_, S_, S_)[lo:hi].reshape(-1, S_, S_)
            if extra2 is not None:
                extra2 = np.asarray(extra2, dtype=self.dtype).reshape(M, T_, S_)[lo:hi].reshape(-1, S_)
            self._surr = (rates, jac2, extra2)
            # mean |J| row-norm per state: sigma_y is identifiable where >> 0 (mcmc.py:240-246)
            self.jac_rownorm = np.sqrt((np.asarray(jac, dtype=np.float64).reshape(-1, S_, S_) ** 2)
                                       .sum(-1)).mean(0)
        spec = model._replace_assignments().sim_input.field_spec()
        self.field = cg.generate_field(spec, sens_theta=True, sens_y0=estimate_initials)
        self._engine = engine
        self._compiled = None
        self._dev_arrays = None
        self._device = device
        self._p2p = None
        self._order = None
        self._evals = 0

    # ------------------------------------------------------------------ likelihood block
    def _local_loglik(self, theta, sigma):
        """[CH, 1+P+S] (log L, dlogL/dtheta, dlogL/dsigma) over THIS rank's series (CUDA)."""
        import torch

        if self._compiled is None:
            self._compiled = self._engine.CompiledModel(self.field, np.dtype(self.dtype).name)
            dev = theta.device
            data, times, y0s, constants = self._arrays
            tt = lambda a: torch.as_tensor(a, device=dev)
            self._dev_arrays = (tt(y0s), tt(constants), tt(times), tt(np.nan_to_num(data)),
                                tt(self.mask.astype(np.uint8)))
        y0s, constants, times, data, mask = self._dev_arrays
        if y0s.shape[0] == 0:
            return torch.zeros((theta.shape[0], 1 + self.field.P + self.field.S),
                               dtype=theta.dtype, device=theta.device)
        if self.mode == "surrogate":
            return self._local_rate_loglik(theta, sigma)
        cfg = self.config
        # A sampler evaluates the same posterior thousands of times with slowly moving parameters:
        # the step counts of an earlier evaluation tell which (chain, series) trajectories are the
        # expensive ones, and handing those out first shortens the critical path of the launch
        # (scheduling hint only: results do not depend on it).  Refreshed every 16 evaluations.
        n_traj = theta.shape[0] * y0s.shape[0]
        order = self._order if (self._order is not None and self._order.numel() == n_traj
                                and self._order.device == theta.device) else None
        out, self._partial, self._status = self._compiled.loglik_grad(
            y0s, theta, constants, times, data, sigma, mask=mask, likelihood=self.likelihood,
            solver=solver_name(cfg.solver), rtol=cfg.rtol, atol=cfg.atol, dt0=cfg.dt0,
            max_steps=cfg.max_steps, order=order)
        if not os.environ.get("CTX_B200_NO_COST_ORDER") and (order is None or self._evals % 16 == 0):
            # (full sort: the likelihood kernel refills lanes straight from global memory, so it
            #  also gains from the lanes of a warp retiring together; measured 1.69 -> 1.28 ms)
            self._order = self._engine.cost_order(*self._compiled.last_loglik_counts, heavy_fraction=1.0)
        self._evals += 1
        return out

    def _local_rate_loglik(self, theta, sigma):
        """Surrogate mode: [CH, 1+P+S] over THIS rank's measurements (ctx_rate_loglik_grad)."""
        import torch

        y0s, constants, times, data, mask = self._dev_arrays
        if getattr(self, "_dev_surr", None) is None:
            dev = theta.device
            rates, jac2, extra2 = self._surr
            tt = lambda a: None if a is None else torch.as_tensor(a, device=dev)
            # evaluation points = the measured states; "{state}_init" = first state of each
            # measurement (mcmc.py:833-842); unobserved (NaN) points are masked out
            Mloc, T, S = data.shape
            pts_ok = mask.reshape(Mloc * T, S).bool().all(dim=1, keepdim=True)
            rmask = (torch.isfinite(tt(rates)) & pts_ok).to(torch.uint8)
            self._dev_surr = (times.reshape(-1).contiguous(), data.reshape(Mloc * T, S).contiguous(),
                              data[:, 0, :].contiguous(), tt(np.nan_to_num(rates)), rmask, tt(jac2),
                              tt(extra2), T)
        t_pts, y_pts, inits, rates, rmask, jac2, extra2, T = self._dev_surr
        return self._compiled.rate_loglik_grad(t_pts, y_pts, theta, constants, inits, rates, jac2,
                                               sigma, n_time=T, mask=rmask, extra2=extra2,
                                               likelihood=self.likelihood)

    def log_likelihood(self, theta, sigma):
        """Likelihood block summed over ALL series (all-reduced when series are sharded)."""
        from .. import parallel

        block = self._local_loglik(theta, sigma)
        if self.shard == "series" and self.world > 1:
            # the path's only collective: the engine's one-shot kernel over NVLink peer memory when
            # the ranks are GPUs of one node (CTX_B200_NO_P2P=1 forces NCCL), torch.distributed else
            if block.is_cuda and not os.environ.get("CTX_B200_NO_P2P"):
                if self._p2p is None or self._p2p.max_bytes < block.numel() * block.element_size():
                    try:
                        self._p2p = parallel.PeerAllReduce(block.numel() * block.element_size())
                    except Exception:
                        self._p2p = False
                if self._p2p:
                    return self._p2p(block)
            parallel.allreduce_sum_(block)
        return block

    # ------------------------------------------------------------------ log joint
    def log_density(self, theta, sigma):
        """theta [CH,P], sigma [CH,S] (torch, on the GPU) ->
        (logp [CH], dlogp/dtheta [CH,P], dlogp/dsigma [CH,S]).

        With a ``shared=True`` error model the sampled site is ONE scalar broadcast over the
        observables (error.py:56-60, 88-91): ``sigma`` is then [CH,1], its prior counts once and
        the returned gradient is [CH,1] (the sum of the per-observable partials)."""
        import torch

        shared = bool(getattr(self.error_model, "shared", False))
        if shared:
            if sigma.shape[-1] != 1:
                raise ValueError("a shared error model samples one scalar: sigma must be [CH, 1]")
            site = sigma
            sigma = site.expand(-1, self.field.S).contiguous()
        block = self.log_likelihood(theta, sigma)
        P = self.field.P
        logp = block[:, 0].clone()
        g_theta = block[:, 1:1 + P].clone()
        g_sigma = block[:, 1 + P:].clone()
        for j, (_name, prior) in enumerate(self.priors):
            logp += prior.log_prob(theta[:, j])
            g_theta[:, j] += prior.grad(theta[:, j])
        hint = torch.as_tensor(self.yerrs, dtype=sigma.dtype, device=sigma.device)
        if shared:
            g_sigma = g_sigma.sum(-1, keepdim=True)
            sigma, hint = site, hint.mean().reshape(1)
        if self.error_model.sampled:
            logp += self.error_model.log_prob(sigma, hint.expand_as(sigma))
            g_sigma += self.error_model.grad(sigma, hint.expand_as(sigma))
        return logp, g_theta, g_sigma

    def __call__(self, *args, **kw):
        """Two call forms.  ``bm(theta, sigma)`` = ``log_density``.  ``bm(y0s, constants, times,
        mask, data)`` -- the reference's model signature (mcmc.py:405: numpyro calls the model with
        the data arguments to trace its log-joint) -- binds those arrays in place of the ones
        extracted at construction and returns the log-density callable ``(theta, sigma) -> (logp,
        dlogp/dtheta, dlogp/dsigma)`` over them."""
        if len(args) + len(kw) == 5:
            names = ("y0s", "constants", "times", "mask", "data")
            vals = dict(zip(names, args))
            vals.update(kw)
            y0s, constants, times, mask, data = (np.asarray(vals[n]) for n in names)
            data = np.where(mask.astype(bool), data, np.nan)     # the kernel's mask is isfinite(data)
            lo, hi = self._series.start, self._series.stop
            self._arrays = tuple(np.asarray(a, dtype=self.dtype)[lo:hi]
                                 for a in (data, times, y0s, constants.reshape(y0s.shape[0], -1)))
            self.mask = np.isfinite(self._arrays[0])
            self._dev_arrays = None
            if self._compiled is not None:
                self._compiled = None
            return self.log_density
        return self.log_density(*args, **kw)
