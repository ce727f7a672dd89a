"""Regularisers added to the training loss (reference: catalax/neural/penalties/).

Same surface as the reference: a penalty is any function ``fun(model, alpha, **kwargs) -> scalar``;
``Penalties`` collects ``Penalty(name, fun, alpha, kwargs)`` items, sums them (penalties.py:84-93),
re-weights them (``update_alpha`` :95-118) and offers the three presets (``for_neural_ode`` :121,
``for_universal_ode`` :145, ``for_rateflow`` :201).  The reference differentiates the penalties with
``jax.grad`` together with the data term; here the data term's gradient comes from the CUDA adjoint
kernels and the penalties -- a few small dense expressions of the parameters -- are differentiated
on the host with torch autograd: a penalty function receives a VIEW of the model whose leaves are
float64 torch tensors and must be written with torch operations.

``model.get_weights_and_biases()`` follows neuralbase.py:576-584: EVERY array leaf of the model --
MLP weights and biases, and also the stoichiometric matrix / mass constraint of a RateFlowODE, the
mechanistic parameters / residual scale / gate matrix of a UniversalODE, and ``max_time`` once
training has replaced it by ``times.max()`` (trainer.py:85-89 makes it a 0-d array) -- so
``l2_regularisation`` shrinks all of them, exactly as the reference does.
"""
from __future__ import annotations

import inspect
from copy import deepcopy
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

import numpy as np


class ModelView:
    """Float64 torch tensors (``requires_grad``) for every array leaf of a neural model."""

    def __init__(self, model):
        import torch

        leaf = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
        self.kind = model.kind
        self.groups: Dict[str, list] = {"weights": [leaf(a) for pair in zip(model.weights, model.biases)
                                                    for a in pair if a is not None]}
        if getattr(model, "rbf_mu", None) is not None:
            self.groups["rbf"] = [leaf(m) for m in model.rbf_mu]
        if getattr(model, "_max_time_is_leaf", False):
            self.groups["max_time"] = [leaf(model.max_time)]
        if self.kind == "rateflow":
            self.stoich_matrix = leaf(model.stoich_matrix)
            self.groups["stoich"] = [self.stoich_matrix]
            self.mass_constraint = None
            if model.mass_constraint is not None:
                self.mass_constraint = leaf(model.mass_constraint)
                self.groups["mass_constraint"] = [self.mass_constraint]
        if self.kind == "universalode":
            self.parameters = leaf(model.parameters)
            self.alpha_residual = leaf(model.alpha_residual)
            self.gate_matrix = leaf(model.gate_matrix)
            self.groups.update(parameters=[self.parameters], alpha=[self.alpha_residual],
                               gate=[self.gate_matrix])

    def get_weights_and_biases(self):
        return [a for g in self.groups.values() for a in g]


# ------------------------------------------------------------------ weight.py
def l2_regularisation(model, alpha: float, **kwargs):
    """alpha * sum of squares of every array leaf (weight.py:6-22)."""
    return alpha * sum((layer ** 2).sum() for layer in model.get_weights_and_biases())


def l1_regularisation(model, alpha: float, **kwargs):
    """alpha * sum of absolute values of every array leaf (weight.py:25-44)."""
    return alpha * sum(layer.abs().sum() for layer in model.get_weights_and_biases())


# ------------------------------------------------------------------ uode.py
def _uode(model):
    assert model.kind == "universalode", "Model must be a UniversalODE"


def l2_reg_gate(model, alpha: float, **kwargs):
    _uode(model)
    return alpha * (model.gate_matrix ** 2).sum()


def l2_reg_alpha(model, alpha: float, **kwargs):
    """On the RAW residual parameter, not on softplus(alpha) (uode.py:33-57)."""
    _uode(model)
    return alpha * (model.alpha_residual ** 2).sum()


def l1_reg_gate(model, alpha: float, **kwargs):
    _uode(model)
    return alpha * model.gate_matrix.abs().sum()


def l1_reg_alpha(model, alpha: float, **kwargs):
    _uode(model)
    return alpha * model.alpha_residual.abs().sum()


# ------------------------------------------------------------------ stoich_mat.py
def _normalize_matrix(stoich_matrix):
    """Columns scaled to unit 2-norm (stoich_mat.py:142-144)."""
    import torch

    return stoich_matrix / torch.linalg.norm(stoich_matrix.abs(), dim=0)


def _rateflow(model):
    assert model.kind == "rateflow", "Model must be a RateFlowODE"
    return _normalize_matrix(model.stoich_matrix)


def penalize_null_space(model, alpha: float = 0.1, **kwargs):
    """alpha * mean(mass_constraint @ S_n); 0 without a constraint (stoich_mat.py:7-26)."""
    import torch

    sn = _rateflow(model)
    if model.mass_constraint is None:
        return torch.zeros((), dtype=torch.float64)
    if model.mass_constraint.shape[1] != sn.shape[0]:
        raise ValueError(
            f"Incompatible shapes for matrix multiplication: mass_constraint.shape="
            f"{tuple(model.mass_constraint.shape)}, stoich_matrix.shape={tuple(sn.shape)}. "
            "Expected mass_constraint.shape[1] == stoich_matrix.shape[0].")
    return alpha * (model.mass_constraint @ sn).mean()


def penalize_density(model, alpha: float = 0.1, **kwargs):
    """alpha * fraction of normalised entries above 0.5 (piecewise constant; stoich_mat.py:29-47)."""
    import torch

    sn = _rateflow(model)
    return alpha * (sn.abs() > 0.5).to(sn.dtype).mean()


def penalize_non_bipolar(model, alpha: float = 0.1, **kwargs):
    """alpha * mean |column sums| (stoich_mat.py:50-71)."""
    return alpha * _rateflow(model).sum(dim=0).abs().mean()


def penalize_non_conservative(model, alpha: float = 0.1, **kwargs):
    """alpha * mean |row sums| (stoich_mat.py:74-91)."""
    return alpha * _rateflow(model).sum(dim=1).abs().mean()


def penalize_duplicate_reactions(model, alpha: float = 1.0, **kwargs):
    """alpha * 0.1 * mean(strict upper triangle of (S_n S_n^T))^2 -- the mean runs over the whole
    matrix, zeros included (stoich_mat.py:94-110)."""
    import torch

    sn = _rateflow(model)
    gram = sn @ sn.T
    n = gram.shape[0]
    upper = torch.triu(torch.ones((n, n), dtype=torch.bool), diagonal=1)
    return alpha * 1e-1 * (torch.where(upper, gram, torch.zeros_like(gram)) ** 2).mean()


def penalize_non_integer(model, alpha: float = 0.1, **kwargs):
    """alpha * mean distance to the nearest integer (stoich_mat.py:113-130)."""
    sn = _rateflow(model)
    return alpha * (sn - sn.round()).abs().mean()


def l1_stoich_penalty(model, alpha: float = 0.1, **kwargs):
    """alpha * induced 1-norm (largest absolute column sum: ``jnp.linalg.norm(S_n, ord=1)`` of a
    matrix; stoich_mat.py:133-139)."""
    import torch

    return alpha * torch.linalg.matrix_norm(_rateflow(model), ord=1)


# ------------------------------------------------------------------ penalties.py
@dataclass
class Penalty:
    """penalties.py:289-383."""
    name: str
    fun: Callable
    alpha: float
    kwargs: Dict[str, Any] = field(default_factory=dict)

    def __post_init__(self):
        if not self.validate_penalty_signature(self.fun):
            raise ValueError(f"Penalty function {getattr(self.fun, '__name__', self.fun)} must accept "
                             "'model' and 'alpha' parameters")

    def __call__(self, model, **kwargs):
        return self.fun(model=model, alpha=self.alpha, **self.kwargs, **kwargs)

    @staticmethod
    def has_model_alpha_signature(func: Callable) -> bool:
        names = list(inspect.signature(func).parameters)
        return "model" in names and "alpha" in names

    @staticmethod
    def validate_penalty_signature(func: Callable) -> bool:
        try:
            inspect.signature(func).bind_partial(model=None, alpha=0.0)
            return True
        except TypeError:
            return False


class Penalties:
    """penalties.py:31-286."""

    def __init__(self, penalties: Optional[List[Penalty]] = None):
        self.penalties = [] if penalties is None else penalties

    def add_penalty(self, name: str, fun: Callable, alpha: float, **kwargs) -> "Penalties":
        self.penalties.append(Penalty(name=name, fun=fun, alpha=alpha, **kwargs))
        return self

    def _total(self, view):
        import torch

        total = torch.zeros((), dtype=torch.float64)
        for penalty in self.penalties:
            total = total + penalty(view)
        return total

    def __call__(self, model) -> float:
        """Sum of all penalties for the model's current parameters."""
        view = model if isinstance(model, ModelView) else ModelView(model)
        return float(self._total(view).detach())

    def value_and_grads(self, model):
        """(value, {leaf group: [gradient arrays]}); groups: ``weights`` (W_0, b_0, W_1, ... as in
        ``zip(weights, biases)`` without the missing biases), ``stoich``, ``gate``, ``alpha``,
        ``parameters`` -- zeros where no penalty touches a leaf."""
        import torch

        view = ModelView(model)
        total = self._total(view)
        leaves = [(g, a) for g, arrs in view.groups.items() for a in arrs]
        grads = {g: [] for g in view.groups}
        if total.requires_grad:
            gs = torch.autograd.grad(total, [a for _, a in leaves], allow_unused=True)
        else:
            gs = [None] * len(leaves)
        for (g, a), d in zip(leaves, gs):
            grads[g].append(np.zeros(tuple(a.shape)) if d is None else d.detach().numpy().copy())
        return float(total.detach()), grads

    def update_alpha(self, alpha: Optional[float], **kwargs) -> "Penalties":
        new = deepcopy(self)
        for penalty in new.penalties:
            if penalty.name in kwargs:
                assert isinstance(kwargs[penalty.name], float), f"Alpha for {penalty.name} must be a float"
                penalty.alpha = kwargs[penalty.name]
            elif alpha is not None:
                penalty.alpha = alpha
        return Penalties(penalties=new.penalties)

    @classmethod
    def for_neural_ode(cls, l2_alpha: Optional[float] = 1e-3, l1_alpha: Optional[float] = None) -> "Penalties":
        p = cls()
        if l2_alpha is not None:
            p.add_penalty(name="l2", fun=l2_regularisation, alpha=l2_alpha)
        if l1_alpha is not None:
            p.add_penalty(name="l1", fun=l1_regularisation, alpha=l1_alpha)
        return p

    @classmethod
    def for_universal_ode(cls, l2_gate_alpha: Optional[float] = 1e-3, l1_gate_alpha: Optional[float] = None,
                          l2_residual_alpha: Optional[float] = 1e-3,
                          l1_residual_alpha: Optional[float] = None,
                          l2_mlp_alpha: Optional[float] = 1e-3,
                          l1_mlp_alpha: Optional[float] = None) -> "Penalties":
        p = cls()
        for name, fun, a in (("l2_gate", l2_reg_gate, l2_gate_alpha), ("l1_gate", l1_reg_gate, l1_gate_alpha),
                             ("l2_residual", l2_reg_alpha, l2_residual_alpha),
                             ("l1_residual", l1_reg_alpha, l1_residual_alpha),
                             ("l2_mlp", l2_regularisation, l2_mlp_alpha),
                             ("l1_mlp", l1_regularisation, l1_mlp_alpha)):
            if a is not None:
                p.add_penalty(name=name, fun=fun, alpha=a)
        return p

    @classmethod
    def for_rateflow(cls, alpha: float = 0.1, density_alpha: Optional[float] = None,
                     bipolar_alpha: Optional[float] = None, conservation_alpha: Optional[float] = None,
                     duplicate_reactions_alpha: Optional[float] = None,
                     integer_alpha: Optional[float] = None, sparsity_alpha: Optional[float] = None,
                     l2_alpha: Optional[float] = None,
                     null_space_alpha: Optional[float] = None) -> "Penalties":
        p = cls()
        p.add_penalty(name="density", fun=penalize_density, alpha=density_alpha or alpha)
        if conservation_alpha is not None:
            p.add_penalty(name="conservation", fun=penalize_non_conservative, alpha=conservation_alpha)
        p.add_penalty(name="bipolar", fun=penalize_non_bipolar, alpha=bipolar_alpha or alpha)
        p.add_penalty(name="duplicate_reactions", fun=penalize_duplicate_reactions,
                      alpha=duplicate_reactions_alpha or alpha)
        p.add_penalty(name="integer", fun=penalize_non_integer, alpha=integer_alpha or alpha)
        p.add_penalty(name="sparsity", fun=l1_stoich_penalty, alpha=sparsity_alpha or alpha)
        p.add_penalty(name="l2", fun=l2_regularisation, alpha=l2_alpha or alpha)
        if null_space_alpha is not None:
            p.add_penalty(name="null_space", fun=penalize_null_space, alpha=null_space_alpha)
        return p


__all__ = ["Penalties", "Penalty", "ModelView", "l2_regularisation", "l1_regularisation",
           "penalize_density", "penalize_non_bipolar", "penalize_non_conservative",
           "penalize_duplicate_reactions", "penalize_non_integer", "penalize_null_space",
           "l1_stoich_penalty", "l2_reg_gate", "l1_reg_gate", "l2_reg_alpha", "l1_reg_alpha"]
