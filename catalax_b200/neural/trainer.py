"""Training curriculum around the CUDA training gradient.

Mirror of ``catalax/neural/strategy.py:16-190`` (``Modes``, ``Step``, ``Strategy.add_step``) and
``catalax/neural/trainer.py:28-253`` (``train_neural_ode``: per-strategy-step optimiser, truncated
time length, shuffled mini-batches, temporal dropout mask, weight / bias scaling, max_time from the
data); the regularisers of ``catalax/neural/penalties`` live in ``penalties.py``.  Every
optimisation step is ONE ``value_and_grad`` of the model = the adjoint kernel pair
(ctx_mlp_adj_forward / backward, Tsit5 or Dopri5); the optimiser update (optax's AdaBelief, the
reference default, or Adam) is element-wise host arithmetic on a few thousand parameters.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Callable, List, Optional

import numpy as np

from .penalties import Penalties, Penalty  # noqa: F401


class Modes(Enum):
    BOTH = "both"
    VECTOR_FIELD = "vector_field"
    MLP = "mlp"


@dataclass
class Step:
    """strategy.py:22-29."""
    steps: int
    batch_size: int
    lr: float = 1e-3
    length: float = 1.0
    penalties: Penalties = field(default_factory=Penalties)
    loss: object = "l2"
    train: Modes = Modes.MLP


class Strategy:
    """strategy.py:119-190: an ordered list of training steps with shared defaults."""

    def __init__(self, penalties: Optional[Penalties] = None, batch_size: int = 5, loss="l2"):
        self.steps: List[Step] = []
        self.penalties = penalties or Penalties()
        self.batch_size = batch_size
        self.loss = loss
        self._print = True

    def add_step(self, lr: float, steps: int, batch_size: Optional[int] = None, length: float = 1.0,
                 penalties: Optional[Penalties] = None, loss=None, train: Modes = Modes.MLP):
        if not (0.0 < length <= 1.0):
            raise ValueError("length must be in (0, 1]")
        self.steps.append(Step(steps=int(steps), batch_size=int(batch_size or self.batch_size), lr=float(lr),
                               length=float(length), penalties=penalties or self.penalties or Penalties(),
                               loss=loss or self.loss, train=Modes(train) if not isinstance(train, Modes) else train))
        return self

    def __iter__(self):
        return iter(self.steps)

    def __len__(self):
        return len(self.steps)


class AdaBelief:
    """optax.adabelief (the reference's default optimiser, trainer.py:33): b1=0.9, b2=0.999,
    eps=1e-16, eps_root=1e-16."""

    def __init__(self, learning_rate: float, b1=0.9, b2=0.999, eps=1e-16, eps_root=1e-16):
        self.lr, self.b1, self.b2, self.eps, self.eps_root = learning_rate, b1, b2, eps, eps_root
        self.t, self.m, self.s = 0, {}, {}

    def update(self, grads: dict) -> dict:
        self.t += 1
        out = {}
        for k, g in grads.items():
            m = self.m.get(k, np.zeros_like(g)); s = self.s.get(k, np.zeros_like(g))
            m = self.b1 * m + (1 - self.b1) * g
            s = self.b2 * s + (1 - self.b2) * (g - m) ** 2 + self.eps_root
            self.m[k], self.s[k] = m, s
            mh, sh = m / (1 - self.b1 ** self.t), s / (1 - self.b2 ** self.t)
            out[k] = -self.lr * mh / (np.sqrt(sh) + self.eps)
        return out


class Adam(AdaBelief):
    """optax.adam: b1=0.9, b2=0.999, eps=1e-8."""

    def __init__(self, learning_rate: float, b1=0.9, b2=0.999, eps=1e-8):
        super().__init__(learning_rate, b1, b2, eps, 0.0)

    def update(self, grads: dict) -> dict:
        self.t += 1
        out = {}
        for k, g in grads.items():
            m = self.b1 * self.m.get(k, np.zeros_like(g)) + (1 - self.b1) * g
            s = self.b2 * self.s.get(k, np.zeros_like(g)) + (1 - self.b2) * g * g
            self.m[k], self.s[k] = m, s
            out[k] = -self.lr * (m / (1 - self.b1 ** self.t)) / (np.sqrt(s / (1 - self.b2 ** self.t)) + self.eps)
        return out


def temporal_dropout_mask(rng: np.random.Generator, n_times: int, temporal_dropout_p: float) -> np.ndarray:
    """trainer.py:256-281: Bernoulli keep-mask over the interior time points; t=0 is always kept."""
    interior = (rng.uniform(size=n_times - 1) < (1.0 - temporal_dropout_p)).astype(np.float64)
    return np.concatenate([np.ones(1), interior])


def dataloader(arrays, batch_size: int, rng: np.random.Generator):
    """trainer.py:409-434: endless shuffled mini-batches; an epoch's ragged tail is dropped."""
    n = arrays[0].shape[0]
    assert all(a.shape[0] == n for a in arrays)
    while True:
        perm = rng.permutation(n)
        start, end = 0, batch_size
        while end < n:
            idx = perm[start:end]
            yield tuple(a[idx] for a in arrays)
            start, end = end, end + batch_size


def _trainable(model, mode: Modes):
    """Which gradient groups a step updates (strategy.py:31-116)."""
    groups = set()
    if mode in (Modes.MLP, Modes.BOTH):
        # everything inside `func` (strategy.py:62): the layers and -- an array leaf since
        # trainer.py:85-89 -- max_time
        groups |= {"weights", "biases", "max_time"}
    if model.kind == "rateflow" and getattr(model, "learn_stoich", False):
        groups.add("stoich")
    if model.kind == "universalode":
        if mode in (Modes.MLP, Modes.BOTH) and model.use_gate:
            groups |= {"gate_matrix", "alpha_residual"}
        if mode in (Modes.VECTOR_FIELD, Modes.BOTH):
            groups.add("parameters")
    return groups


def train_neural_ode(model, dataset, strategy: Strategy, optimizer: Callable = AdaBelief,
                     validation_dataset=None, print_every: int = 10, weight_scale: float = 1e-8,
                     solver=None, log: Optional[str] = None, seed: int = 420, progress_bar: bool = False,
                     temporal_dropout_p: float = 0.0):
    """trainer.py:28-253.  Trains ``model`` IN PLACE on ``dataset`` following ``strategy`` and
    returns it together with nothing else (the reference returns the updated pytree)."""
    data, times, y0s = dataset.to_jax_arrays(model.state_order, inits_to_array=True)
    data, times, y0s = (np.asarray(a, dtype=np.float64) for a in (data, times, y0s))
    rng = np.random.default_rng(seed)
    n_times = data.shape[1]
    # _scale_weights (trainer.py:459-495): every weight and every hidden bias times weight_scale
    model.weights = [w * weight_scale for w in model.weights]
    model.biases = [None if b is None else (b * weight_scale if l < len(model.biases) - 1 else b)
                    for l, b in enumerate(model.biases)]
    model.max_time = float(times.max())
    # trainer.py:85-89 replaces max_time by the ARRAY times.max(): from here on it is one of the
    # leaves get_weights_and_biases() returns (so l1 / l2 regularisation count it) and, sitting
    # inside `func` (strategy.py:62), one of the parameters the optimiser moves.
    model._max_time_is_leaf = True
    model._compiled.clear()
    if solver is not None:
        model.solver = solver
    history = []
    if log is not None:
        with open(log, "w") as f:
            f.write("strategy\tstep\tmean_loss\tmae\n")
    for s_idx, st in enumerate(strategy):
        assert st.batch_size < data.shape[0], (
            f"Batch size of strategy #{s_idx} ({st.batch_size}) is larger than the dataset size "
            f"({data.shape[0]}). Please reduce the batch size to be < {data.shape[0]}.")
        opt = optimizer(learning_rate=st.lr)
        groups = _trainable(model, st.train)
        n_keep = max(int(n_times * st.length) + 1, 2)
        _t, _d = times[:, :n_keep], data[:, :n_keep, :]
        batches = dataloader((_d, y0s, _t), st.batch_size, rng)
        for step, (yi, y0i, ti) in zip(range(st.steps), batches):
            mask = temporal_dropout_mask(rng, ti.shape[1], temporal_dropout_p)
            val, g = model.value_and_grad(ti, y0i, yi, loss=st.loss, mask=mask)
            pval, pg = st.penalties.value_and_grads(model)
            flat = {}
            if "weights" in groups:
                k = 0
                for l, (gw, gb) in enumerate(zip(g["weights"], g["biases"])):
                    flat[f"W{l}"] = gw + pg["weights"][k]; k += 1
                    if gb is not None:
                        flat[f"b{l}"] = gb + pg["weights"][k]; k += 1
            if "max_time" in groups:
                flat["max_time"] = np.asarray(g["max_time"] + float(pg["max_time"][0]))
            if "stoich" in groups:
                flat["stoich"] = g["stoich"] + pg["stoich"][0]
            if "gate_matrix" in groups:
                flat["gate"] = g["gate_matrix"] + pg["gate"][0]
                flat["alpha"] = g["alpha_residual"] + pg["alpha"][0]
            if "parameters" in groups:
                flat["parameters"] = g["parameters"] + pg["parameters"][0]
            upd = opt.update(flat)
            L = len(model.weights)
            zeros_w = [np.zeros_like(w) for w in model.weights]
            zeros_b = [None if b is None else np.zeros_like(b) for b in model.biases]
            model.apply_updates([upd.get(f"W{l}", zeros_w[l]) for l in range(L)],
                                [None if model.biases[l] is None else upd.get(f"b{l}", zeros_b[l]) for l in range(L)],
                                dstoich=upd.get("stoich"), dgate=upd.get("gate"), dalpha=upd.get("alpha"),
                                dparameters=upd.get("parameters"),
                                dmax_time=None if "max_time" not in upd else float(upd["max_time"]))
            history.append(val + pval)
            if (step % print_every) == 0 or step == st.steps - 1:
                pred = model.predict_arrays(_t, y0s)
                mean_loss = float(np.mean(0.5 * (pred - _d) ** 2))
                mae = float(np.mean(np.abs(pred - _d)))
                if progress_bar and strategy._print:
                    print(f"step #{s_idx + 1} {step}: loss {mean_loss:.4f} mae {mae:.4f}", flush=True)
                if log is not None:
                    with open(log, "a") as f:
                        f.write(f"{s_idx}\t{step}\t{mean_loss}\t{mae}\n")
    model.history = history
    return model
