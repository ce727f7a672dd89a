"""Parameter fitting by repeated simulation: the reference's ``optimize`` residual loop.

Mirror of ``catalax/tools/optimization.py``:

* ``optimize`` (:14-99): bounds / initial values from the model's parameters
  (``_initialize_params`` :122-165), data and time grids from the dataset, ``dataset.to_config()``
  with ``dt0`` / ``max_steps`` overridden, then a minimiser over the parameter vector;
* ``_residual`` (:168-207): ``model.simulate(dataset, config, saveat=times, parameters=theta)``
  followed by ``objective_fun(data, states[:, :, observables])`` (default ``optax.l2_loss`` =
  ``0.5 (pred - target)^2``, element-wise).

The reference hands the residual ARRAY to lmfit (un-vendored dependency, not in this image);
lmfit's scalar methods (its default here is "bfgs") minimise ``sum(residual**2)``.  The same
objective is minimised here with scipy (L-BFGS-B for the scalar methods, trust-region least
squares for "leastsq" / "least_squares"), and -- unlike the reference, which lets lmfit
difference the solve numerically -- with the EXACT gradient of the discrete solve: every
evaluation is one ``ctx_solve_sens`` launch (forward tangents w.r.t. the parameters, the same
kernel the NUTS path uses), so one objective + gradient costs one batched solve instead of
``2 P + 1``.  Entries whose datum is NaN (padded measurements) do not contribute.
"""
from __future__ import annotations

from copy import deepcopy
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Tuple

import numpy as np

from ..config import default_dtype
from ..model.inaxes import InAxes  # noqa: F401  (part of the mirrored surface)
from .simulation import SolveError


def l2_loss(predictions, targets):
    """``optax.l2_loss``: 0.5 * (predictions - targets)**2, element-wise (torch or numpy)."""
    return 0.5 * (predictions - targets) ** 2


@dataclass
class MinimizerResult:
    """The fields of lmfit's ``MinimizerResult`` the reference's callers read."""

    params: Dict[str, float]
    success: bool
    message: str
    nfev: int
    chisqr: float                       # sum(residual**2) at the optimum
    residual: np.ndarray
    method: str
    init_vals: Dict[str, float] = field(default_factory=dict)


def _initialize_params(model, global_upper_bound, global_lower_bound):
    """name -> (value, lower, upper, vary) with the reference's rules (optimization.py:122-165)."""
    out = {}
    for param in model.parameters.values():
        if param.lower_bound is None:
            assert global_lower_bound is not None, (
                f"Neither a global lower bound nor lower bound for parameter '{param.name}' is given.")
            param.lower_bound = global_lower_bound
        if param.upper_bound is None:
            assert global_upper_bound is not None, (
                f"Neither a global upper bound nor upper bound for parameter '{param.name}' is given.")
            param.upper_bound = global_upper_bound
        if param.initial_value is None and param.value is None:
            raise ValueError(
                f"Neither 'value' nor 'initial_value' given for parameter '{param.name}'. "
                "Please add an initial value for the optimization to work.")
        elif param.initial_value is None and param.value is not None:
            param.initial_value = param.value
        out[param.name] = (float(param.initial_value), float(param.lower_bound),
                           float(param.upper_bound), not param.constant)
    return out


class _Objective:
    """theta -> (sum(residual**2), gradient, residual, d residual / d theta) on the GPU."""

    def __init__(self, model, dataset, dt0, max_steps, objective_fun):
        import torch

        from .simulation import Simulation

        self.torch = torch
        order = model.get_state_order()
        obs_names = dataset.get_observable_states_order()
        self.obs = [i for i, s in enumerate(order) if s in obs_names]
        data, times, _ = dataset.to_jax_arrays(order)
        config = dataset.to_config()
        config.dt0 = dt0
        config.max_steps = max_steps
        config.nsteps = None
        self.dtype = default_dtype()
        m = model._replace_assignments()

        class _Sens:                    # stands in for InAxes-like ``sensitivity`` (index of 0 = argnum)
            value = (None, 0, None, None)

        self.f, _ = Simulation(m.sim_input, config, sensitivity=_Sens())._prepare_func(
            in_axes=(0, None, 0, 0))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        tt = lambda a: torch.as_tensor(np.asarray(a, dtype=self.dtype), device=dev)
        self.y0 = tt(dataset.to_y0_matrix(state_order=order))
        self.consts = tt(dataset.to_y0_matrix(state_order=m.get_constants_order()))
        self.times = tt(times)
        d = np.asarray(data, dtype=np.float64)[:, :, self.obs]
        self.mask = torch.as_tensor(np.isfinite(d), device=dev)
        self.data = torch.as_tensor(np.nan_to_num(d), device=dev)
        self.objective_fun = objective_fun
        self.nfev = 0

    def __call__(self, theta: np.ndarray):
        torch = self.torch
        self.nfev += 1
        th = torch.as_tensor(np.asarray(theta, dtype=self.dtype), device=self.dev)
        sens = self.f(self.y0, th, self.consts, self.times)          # [M,T,S,P]
        ys = self.f.last.ys                                          # [M,T,S]
        pred = ys[:, :, self.obs].to(torch.float64).detach().requires_grad_(True)
        r = self.objective_fun(pred, self.data)
        r = torch.where(self.mask, r, torch.zeros_like(r))
        (dr_dpred,) = torch.autograd.grad(r.sum(), pred, retain_graph=False)   # element-wise objective
        dpred = sens[:, :, self.obs, :].to(torch.float64)            # [M,T,n_obs,P]
        jac = (torch.where(self.mask, dr_dpred, torch.zeros_like(dr_dpred))[..., None] * dpred)
        r = r.detach()
        F = float((r * r).sum().item())
        g = (2.0 * r[..., None] * jac).sum(dim=(0, 1, 2)).cpu().numpy()
        return F, g, r.reshape(-1).cpu().numpy(), jac.reshape(-1, jac.shape[-1]).cpu().numpy()


def optimize(model, dataset, global_upper_bound: Optional[float] = 1e5,
             global_lower_bound: Optional[float] = 1e-6, dt0: float = 0.01,
             objective_fun: Callable = l2_loss, max_steps: int = 64 ** 4,
             method: str = "bfgs") -> Tuple[MinimizerResult, "object"]:
    """Fit the model's parameters to a dataset (optimization.py:14-99).

    Returns ``(result, new_model)``; ``new_model`` is a modified COPY carrying the inferred values.
    ``objective_fun(predictions, targets)`` must be element-wise and written with operators torch
    understands (the default is ``l2_loss``)."""
    from scipy import optimize as so

    spec = _initialize_params(model, global_upper_bound, global_lower_bound)
    names = model.get_parameter_order()
    x0 = np.array([spec[n][0] for n in names], dtype=np.float64)
    lo = np.array([spec[n][1] for n in names], dtype=np.float64)
    hi = np.array([spec[n][2] for n in names], dtype=np.float64)
    vary = np.array([spec[n][3] for n in names], dtype=bool)
    obj = _Objective(model, dataset, dt0, max_steps, objective_fun)

    def full(xv):
        x = x0.copy()
        x[vary] = xv
        return x

    last = {}
    if method.lower() in ("leastsq", "least_squares"):
        def res(xv):
            F, g, r, J = obj(full(xv))
            last.update(F=F, r=r, J=J[:, vary])
            return r

        sol = so.least_squares(res, x0[vary], jac=lambda xv: last["J"], bounds=(lo[vary], hi[vary]),
                               method="trf", x_scale="jac")
        xv, ok, msg = sol.x, bool(sol.success), str(sol.message)
    else:
        def fg(xv):
            try:
                F, g, r, _ = obj(full(xv))
            except SolveError:
                # a trial point the solver cannot integrate within max_steps (the reference
                # propagates diffrax's exception out of lmfit): reject it, the line search backtracks
                return 1e300, np.zeros(int(vary.sum()))
            last.update(F=F, r=r)
            return F, g[vary]

        sol = so.minimize(fg, x0[vary], jac=True, method="L-BFGS-B",
                          bounds=list(zip(lo[vary], hi[vary])),
                          options=dict(maxiter=500, ftol=1e-15, gtol=1e-12))
        xv, ok, msg = sol.x, bool(sol.success), str(sol.message)
    F, _, r, _ = obj(full(xv))
    values = dict(zip(names, full(xv).tolist()))
    result = MinimizerResult(params=values, success=ok, message=msg, nfev=obj.nfev, chisqr=F,
                             residual=r, method=method, init_vals=dict(zip(names, x0.tolist())))
    new_model = deepcopy(model)
    for name, value in values.items():
        new_model.parameters[name].value = value
    return result, new_model
