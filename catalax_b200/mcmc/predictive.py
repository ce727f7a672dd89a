"""Posterior re-integration: one solve per posterior draw x measurement series.

Mirror of ``catalax/mcmc/predictive.py:109-155`` (``_integrate_ensemble``:
``jax.vmap(per_draw)(thetas)`` over ``sim_func(y0s, theta, constants, times)`` ->
``[n_draws, n_meas, n_steps, n_obs]``) and of the mechanistic point-wise log-likelihood that
LOO uses (``catalax/mcmc/loo.py:131-200``).  Both are the same CUDA kernel as ``Model.simulate``
with a two-level batch (``ctx_solve_cfg.inner``): parameters are indexed by draw, initial values
/ constants / times by series, so nothing is tiled or copied on the host.
"""
from __future__ import annotations

import math

import numpy as np

from ..config import default_dtype
from ..solvers import solver_name
from ..tools import codegen as cg


def ensemble_times(times: np.ndarray, n_steps: int) -> np.ndarray:
    """Per-series dense grid t0 + (t1 - t0) * linspace(0, 1, n_steps) (predictive.py:140-143)."""
    times = np.asarray(times)
    t0, t1 = times.min(axis=-1), times.max(axis=-1)
    frac = np.linspace(0.0, 1.0, n_steps)
    return (t0[:, None] + (t1 - t0)[:, None] * frac[None, :]).astype(times.dtype)


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

    def integrate(self, y0s, thetas, constants, times, device=None):
        """y0s [M,S], thetas [D,P], constants [M,C], times [M,T] -> torch CUDA [D,M,T,S]."""
        import torch

        dev = device or torch.device("cuda", torch.cuda.current_device())
        tt = lambda a: (a if isinstance(a, torch.Tensor) else
                        torch.as_tensor(np.asarray(a, dtype=self.dtype))).to(dev)
        y0s, thetas, constants, times = tt(y0s), tt(thetas), tt(constants), tt(times)
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
                                 likelihood="normal"):
        """log p(data[m,t,o] | draw d) for every draw and observation -> torch [D,M,T,S]
        (NaN where the observation is missing).  The input of LOO's PSIS (loo.py:131-200)."""
        import torch

        ys = self.integrate(y0s, thetas, constants, times)
        dev = ys.device
        data_t = torch.as_tensor(np.asarray(data, dtype=self.dtype), device=dev)
        s = torch.as_tensor(np.asarray(sigmas, dtype=self.dtype), device=dev).abs()[:, None, None, :]
        z = (data_t[None] - ys) / s
        if likelihood == "normal":
            lp = -0.5 * z * z - torch.log(s) - 0.5 * math.log(2 * math.pi)
        elif likelihood == "softlaplace":
            lp = math.log(2 / math.pi) - torch.log(s) - torch.logaddexp(z, -z)
        else:
            raise NotImplementedError(likelihood)
        return lp
