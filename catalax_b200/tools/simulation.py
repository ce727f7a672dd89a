"""The solve seam: ``Simulation._prepare_func(in_axes) -> (f, stack)`` with
``f(y0, parameters, constants, time) -> ys``.

Mirror of ``catalax/tools/simulation.py`` (SimulationInput :31-108, Simulation :111-355).
Where the reference builds ``eqx.filter_jit(jax.vmap(diffeqsolve(...)))`` (:236-275, :346-355)
this builds a :class:`SolveFunction` bound to NVRTC-compiled sm_100a kernels.  Shapes follow
the 4-tuple ``in_axes`` of ``0 | None`` (inaxes.py:5-9); dtype follows
``catalax_b200.enable_x64`` like JAX's x64 switch.  Inputs may be numpy arrays (host: copied
to the GPU, result returned as numpy) or torch CUDA tensors (device-resident: result is a torch
CUDA tensor, nothing crosses PCIe).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Tuple

import numpy as np

from ..config import default_dtype
from ..solvers import NON_ADAPTIVE_SOLVERS, solver_name
from . import codegen as cg


class SolveError(RuntimeError):
    """Raised when ``throw=True`` and a trajectory failed (diffrax raises in the reference)."""


@dataclass
class SimulationInput:
    """Equations + ordered symbol lists (simulation.py:31-108)."""

    states: List[str]
    parameters: List[str]
    constants: List[str]
    odes: List[Any] = field(default_factory=list)
    reactions: List[Any] = field(default_factory=list)

    @property
    def type(self) -> str:
        if len(self.reactions) > 0 and len(self.odes) == 0:
            return "reaction"
        if len(self.reactions) == 0 and len(self.odes) > 0:
            return "ode"
        return "mixed"

    def field_spec(self) -> cg.FieldSpec:
        return cg.FieldSpec(
            states=list(self.states), parameters=list(self.parameters),
            constants=list(self.constants),
            odes={str(o.state.symbol): o.equation for o in self.odes},
            reactions=[(r.symbol, [(e.stoichiometry, e.state) for e in r.reactants],
                        [(e.stoichiometry, e.state) for e in r.products], r.equation)
                       for r in self.reactions])

    @property
    def stack(self) -> cg.GeneratedField:
        """The generated vector field plays the role of ODEStack/ReactionStack/MixedStack."""
        return cg.generate_field(self.field_spec())


class SolveFunction:
    """Callable with the reference's signature ``f(y0, parameters, constants, time) -> ys``."""

    def __init__(self, sim_input: SimulationInput, config, in_axes: Optional[Tuple], dtype,
                 sensitivity_index: Optional[int] = None):
        from .. import engine

        self.sim_input = sim_input
        self.in_axes = in_axes
        self.dtype = np.dtype(dtype)
        self.solver = solver_name(config.solver)
        self.rtol, self.atol, self.dt0 = config.rtol, config.atol, config.dt0
        self.max_steps, self.throw = config.max_steps, config.throw
        self.sensitivity_index = sensitivity_index
        if sensitivity_index not in (None, 0, 1):
            raise NotImplementedError("sensitivities are available w.r.t. y0 (0) or parameters (1)")
        self.field = cg.generate_field(sim_input.field_spec(),
                                       sens_theta=sensitivity_index == 1,
                                       sens_y0=sensitivity_index == 0)
        self.model = engine.CompiledModel(self.field, self.dtype.name)
        self.last = None  # SolveOutput of the most recent call (status, step counts)
        self._order = None  # cost order of an earlier call with the same batch size (scheduling hint)
        self._calls = 0

    def __call__(self, y0, parameters, constants, time):
        import torch

        from .. import engine

        axes = self.in_axes if self.in_axes is not None else (None, None, None, None)
        args = [y0, parameters, constants, time]
        on_device = any(isinstance(a, torch.Tensor) and a.is_cuda for a in args)
        if not torch.cuda.is_available():
            raise engine.EngineError("no CUDA device visible: catalax_b200 has no CPU fallback")
        dev = next((a.device for a in args if isinstance(a, torch.Tensor) and a.is_cuda),
                   torch.device("cuda", torch.cuda.current_device()))
        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        widths = [self.field.S, self.field.P, self.field.C, None]
        dev_args = []
        for a, ax, w in zip(args, axes, widths):
            if not isinstance(a, torch.Tensor):
                a = torch.as_tensor(np.asarray(a, dtype=self.dtype))
            a = a.to(device=dev, dtype=tdt)
            if w == 0:
                a = a.reshape((a.shape[0], 0) if ax == 0 else (0,))
            dev_args.append(a)
        # Repeated solves of the same batch (an optimiser's objective, a sweep over parameters) have
        # similar per-row costs: the step counts of an earlier call order the rows of this one,
        # most expensive first (scheduling hint only -- results do not depend on it; refreshed
        # every 8 calls; large batches only).
        Bs = [a.shape[0] for a, ax in zip(dev_args, axes) if ax == 0]
        B = Bs[0] if Bs else 1
        order = self._order if (self._order is not None and self._order.numel() == B
                                and self._order.device == dev) else None
        out = self.model.solve(*dev_args, solver=self.solver, in_axes=axes, rtol=self.rtol,
                               atol=self.atol, dt0=self.dt0, max_steps=self.max_steps,
                               tangents=self.sensitivity_index is not None, order=order)
        self._calls += 1
        if B >= 4096 and (order is None or self._calls % 8 == 0):
            self._order = out.cost_order()
        self.last = out
        if self.throw:
            bad = int((out.status != 0).sum().item())
            if bad:
                code = int(out.status[out.status != 0][0].item())
                why = {1: "the maximum number of solver steps was reached",
                       2: "the step size became non-finite"}.get(code, f"status {code}")
                raise SolveError(f"{bad} of {out.status.numel()} trajectories failed: {why} "
                                 f"(max_steps={self.max_steps}); pass throw=False to get inf instead")
        res = out.sens if self.sensitivity_index is not None else out.ys
        if all(ax is None for ax in axes):
            res = res[0]
        return res if on_device else res.cpu().numpy()


class Simulation:
    """Builds the solve function for a system of ODEs (simulation.py:111-355)."""

    def __init__(self, sim_input: SimulationInput, config, sensitivity=None):
        self.sim_input = sim_input
        self.config = config
        self.sensitivity = sensitivity
        self._simulation_func: Optional[Callable] = None

    def __call__(self, y0, parameters, constants, time):
        if self._simulation_func is None:
            raise ValueError("Simulation function not initialized")
        return self._simulation_func(y0, parameters, constants, time)

    def _create_controller(self) -> str:
        """'constant' for Euler/Heun, 'pid' otherwise (simulation.py:198-212)."""
        solver = self.config.solver
        cls = solver if isinstance(solver, type) else type(solver)
        return "constant" if cls in NON_ADAPTIVE_SOLVERS else "pid"

    def _prepare_func(self, in_axes: Optional[Tuple] = None):
        sens_index = None
        if self.sensitivity is not None:
            sens_index = self.sensitivity.value.index(0)  # simulation.py:311
        f = SolveFunction(self.sim_input, self.config, in_axes, default_dtype(), sens_index)
        self._simulation_func = f
        return f, self
