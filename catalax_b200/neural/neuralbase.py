"""NeuralODE / RateFlowODE / UniversalODE forward path on the B200 engine.

Mirror of the forward (predict / rates) surface of ``catalax/neural``:
``NeuralBase.__call__`` / ``predict`` (neuralbase.py:134-143, 252-317), ``rates`` (:319-327),
the per-class solve calls (neuralode.py:38-59, rateflow.py:100-121, universalode.py:118-141) and
the controller defaults (neuralbase.py:623-645: rtol=1e-3, atol=1e-6), the ``.eqx`` checkpoint
format of all three classes (neuralbase.py:481-571) and the training gradient (trainer.py:268-306).

Weights follow equinox's convention (``Linear(x) = W @ x + b``, ``W: [out, in]``); they are given
explicitly or drawn from ``numpy.random.default_rng(seed)`` with ``N(0, 1/fan_in)``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from ..config import default_dtype
from ..solvers import Tsit5, solver_name
from ..tools import codegen_mlp as cm


def _act_name(a) -> str:
    if a is None:
        return "identity"
    name = a if isinstance(a, str) else getattr(a, "__name__", str(a))
    name = name.lower()
    if name not in cm.ACTIVATIONS:
        raise NotImplementedError(f"activation {a!r}; supported: {sorted(cm.ACTIVATIONS)}")
    return name


class RBFLayer:
    """Stand-in for ``catalax.neural.RBFLayer(gamma)`` (rbf.py:10-39) as an ``activation=``
    argument: every hidden activation becomes ``exp(-gamma * sum_j (x_i - mu_j)^2)`` with learned
    centres ``mu`` [width] per hidden layer (mlp.py:86-118)."""

    def __init__(self, gamma: float, width_size: Optional[int] = None, *, key=None):
        self.gamma = float(gamma)


def masked_loss(pred, target, loss="l2", mask=None):
    """The training loss of ``grad_loss`` (trainer.py:270-311) without the penalties, on torch
    tensors: per-point loss ``[B,T,S]`` weighted by the temporal-dropout mask ``[T]`` and normalised
    by the KEPT points, ``sum(mask * per_point) / (sum(mask) * B * S)``.  ``loss``: "l2"
    (``optax.l2_loss`` = 0.5 (p - y)^2), "l1", or a callable ``(pred, target) -> per-point loss``."""
    import torch

    if callable(loss):
        per_point = loss(pred, target)
    elif loss == "l2":
        per_point = 0.5 * (pred - target) ** 2
    elif loss == "l1":
        per_point = (pred - target).abs()
    else:
        raise ValueError(f"unknown loss {loss!r}")
    B, T, S = pred.shape
    m = torch.ones(T, dtype=pred.dtype, device=pred.device) if mask is None else mask.to(pred.dtype)
    return (per_point * m[None, :, None]).sum() / (m.sum() * B * S)


class NeuralBase:
    kind = "neuralode"
    default_activation = "tanh"      # neuralbase.py:79 (jax.nn.tanh)

    def __init__(self, data_size: int, width_size: int, depth: int, state_order: List[str],
                 observable_indices: Optional[List[int]] = None, solver=Tsit5,
                 activation="tanh", final_activation=None, use_final_bias: bool = False,
                 out_size: Optional[int] = None, max_time: float = 1.0, *, key=0,
                 weights: Optional[Sequence[np.ndarray]] = None,
                 biases: Optional[Sequence[Optional[np.ndarray]]] = None, mech=None,
                 rbf_mu: Optional[Sequence[np.ndarray]] = None, **hyper):
        # activation=RBFLayer(gamma) replaces every hidden activation by a radial-basis layer with
        # its own centres mu [width] (mlp.py:86-118, rbf.py:14-39)
        rbf = isinstance(activation, RBFLayer)
        self.rbf_gamma = float(activation.gamma) if rbf else None
        if rbf:
            activation = "identity"
        self.hyperparams = dict(hyper)     # extra constructor kwargs travel in the checkpoint header
        self.data_size, self.width_size, self.depth = data_size, width_size, depth
        self.state_order = list(state_order)
        self.observable_indices = list(observable_indices) if observable_indices is not None \
            else list(range(data_size))
        self.solver = solver
        self.max_time = float(max_time)
        self.spec = cm.MLPSpec(kind=self.kind, data_size=data_size, width_size=width_size,
                               depth=depth, out_size=out_size or data_size,
                               activation=_act_name(activation),
                               final_activation=_act_name(final_activation),
                               use_final_bias=use_final_bias, mech=mech, rbf=rbf)
        rng = np.random.default_rng(key)
        dims = self.spec.layer_dims()
        if weights is None:
            weights = [rng.normal(0.0, 1.0 / np.sqrt(i), (o, i)) for i, o in dims]
        if biases is None:
            biases = [rng.normal(0.0, 1.0 / np.sqrt(i), (o,)) for i, o in dims]
            if not use_final_bias:
                biases[-1] = None
        self.weights = [np.asarray(w, dtype=np.float64) for w in weights]
        self.biases = [None if b is None else np.asarray(b, dtype=np.float64) for b in biases]
        self.rbf_mu = None
        if rbf:     # rbf.py:27: centres ~ U(0, 1), one vector per hidden layer
            self.rbf_mu = [np.asarray(m, dtype=np.float64) for m in rbf_mu] if rbf_mu is not None \
                else [rng.uniform(0.0, 1.0, (width_size,)) for _ in range(depth)]
        self._compiled = {}
        self._rng = rng

    # ------------------------------------------------------------------ engine plumbing
    def _extra(self):
        return {}

    def _pack_extra(self):
        e = dict(self._extra())
        if self.spec.rbf:
            e.update(rbf_mu=self.rbf_mu, rbf_gamma=self.rbf_gamma)
        return e

    def _t0_from_ts(self) -> bool:
        return True     # neuralode.py:52 / rateflow.py:114: t0 = ts[0]

    def _max_steps(self) -> int:
        return 4096     # diffrax's default max_steps (the solve calls pass none)

    def _model(self, dtype):
        from .. import engine

        key = np.dtype(dtype).name
        if key not in self._compiled:
            field = cm.generate_mlp_field(self.spec)
            theta = cm.pack_parameters(self.spec, self.weights, self.biases, self.max_time,
                                       np.dtype(dtype), **self._pack_extra())
            self._compiled[key] = (engine.CompiledModel(field, key), theta)
        return self._compiled[key]

    def tensor_core_eligible(self, dtype) -> bool:
        """f32 models whose padded shapes fit tcgen05.mma tiles run on the tensor cores."""
        return np.dtype(dtype) == np.float32 and cm.tc_dims(self.spec) is not None

    def _tc_theta(self):
        if "tc" not in self._compiled:
            self._compiled["tc"] = cm.pack_tc_parameters(self.spec, self.weights, self.biases,
                                                         self.max_time, **self._extra())
        return self._compiled["tc"]

    # ------------------------------------------------------------------ reference checkpoints
    # File layout written by the reference's ``save_to_eqx`` (neuralbase.py:530-571): one JSON
    # line ``{"hyperparameters": {...}}`` (the constructor arguments), then
    # ``eqx.tree_serialise_leaves``: every leaf of the module pytree that is an array or a Python
    # / numpy scalar as consecutive ``numpy.save`` blobs, in field order --
    #   func.mlp.layers: W_0, b_0, [mu_0, gamma_0 (rbf)], ..., W_last[, b_last]; func.max_time;
    #   observable_indices (ints); hyperparams (numeric leaves, keys sorted); then the subclass
    #   fields: RateFlowODE stoich_matrix [S,R], reaction_size, mass_constraint [n_c,S] | absent
    #   (rateflow.py:19-27); UniversalODE parameters [P], alpha_residual [S], gate_matrix [S,S]
    #   (universalode.py:20-26; its ``model`` dict and ``use_gate`` are static).
    # Strings, None and static fields are not leaves.  Verified on the checkpoint the reference
    # ships (examples/trained/menten_trained.eqx: 5 arrays, max_time, then 7 scalar leaves).
    @staticmethod
    def _read_eqx(path):
        import json

        with open(path, "rb") as f:
            header = json.loads(f.readline().decode())
            leaves = []
            while True:
                try:
                    leaves.append(np.load(f, allow_pickle=False))
                except (EOFError, ValueError):
                    break
        return dict(header["hyperparameters"]), leaves

    @staticmethod
    def _mlp_from_leaves(path, hp, leaves):
        S, W, depth = int(hp["data_size"]), int(hp["width_size"]), int(hp["depth"])
        out = int(hp["out_size"]) if hp.get("out_size") else S
        rbf = bool(hp.get("rbf"))
        dims = [S + 1] + [W] * depth + [out]
        weights, biases, mus, gammas, k = [], [], [], [], 0
        for l, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
            Wm = leaves[k]; k += 1
            if Wm.shape != (o, i):
                raise ValueError(f"{path}: layer {l} weight has shape {Wm.shape}, expected {(o, i)}")
            weights.append(Wm)
            if l < depth or bool(hp.get("use_final_bias")):
                biases.append(leaves[k]); k += 1
            else:
                biases.append(None)
            if rbf and l < depth:
                mu = leaves[k]; k += 1
                if mu.shape != (W,):
                    raise ValueError(f"{path}: RBF centres of layer {l} have shape {mu.shape}")
                mus.append(mu)
                gammas.append(float(leaves[k])); k += 1
        max_time = float(leaves[k]) if k < len(leaves) and leaves[k].shape == () else 1.0
        k += 1
        return weights, biases, mus, gammas, max_time, k

    @classmethod
    def from_eqx(cls, path, **overrides):
        """Load a model saved by the reference's ``save_to_eqx`` (neuralbase.py:481-571):
        NeuralODE (also with RBF hidden layers), RateFlowODE and UniversalODE checkpoints."""
        hp, leaves = cls._read_eqx(path)
        weights, biases, mus, gammas, max_time, k = cls._mlp_from_leaves(path, hp, leaves)
        S = int(hp["data_size"])
        rest = [a for a in leaves[k:] if a.ndim > 0]      # the scalar leaves repeat the header
        if mus and len(set(gammas)) > 1:
            raise NotImplementedError("RBF layers with different gammas")
        # neuralbase.py:508-518: a header without the key falls back to the class default
        activation = RBFLayer(gammas[0]) if mus else hp.get("activation", cls.default_activation)
        kw = dict(data_size=S, width_size=int(hp["width_size"]), depth=int(hp["depth"]),
                  observable_indices=hp.get("observable_indices", [0]),
                  activation=activation, use_final_bias=bool(hp.get("use_final_bias")),
                  max_time=max_time, weights=weights, biases=biases)
        if mus:
            kw["rbf_mu"] = mus
        if cls.kind == "neuralode":
            kw.update(state_order=hp["state_order"], final_activation=hp.get("final_activation"),
                      out_size=hp.get("out_size"))
        elif cls.kind == "rateflow":
            R_ = int(hp.get("reaction_size") or hp["out_size"])
            sm = next((a for a in rest if a.shape == (S, R_)), None)
            if sm is None:
                raise ValueError(f"{path}: no stoichiometric matrix of shape {(S, R_)} among the leaves")
            after = rest[next(i for i, a in enumerate(rest) if a is sm) + 1:]
            mc = next((a for a in after if a.ndim == 2 and a.shape[1] == S), None)
            kw.update(state_order=hp["state_order"], reaction_size=R_, stoich_matrix=sm,
                      learn_stoich=bool(hp.get("learn_stoich", True)), mass_constraint=mc)
        else:   # universalode: the mechanistic model travels in the header (universalode.py:63)
            from ..model import Model

            md = hp.get("model")
            if md is None:
                raise ValueError(f"{path}: UniversalODE checkpoint without a model in its header")
            if isinstance(md, str):
                import json

                md = json.loads(md)
            model = Model.from_dict(md)
            if len(rest) < 3 or rest[-1].shape != (S, S) or rest[-2].shape != (S,):
                raise ValueError(f"{path}: expected parameters, alpha_residual [S], gate_matrix [S,S] "
                                 "as the last array leaves")
            kw.update(model=model, final_activation=hp.get("final_activation") or "identity",
                      gate_matrix=rest[-1], alpha_residual=rest[-2], parameters=rest[-3])
        kw.update(overrides)
        return cls(**kw)

    def _header(self):
        hp = {"data_size": self.data_size, "width_size": self.width_size, "depth": self.depth,
              "out_size": None if self.kind != "rateflow" else self.spec.out_size,
              "rbf": bool(self.spec.rbf), "use_final_bias": bool(self.spec.use_final_bias),
              "observable_indices": list(self.observable_indices),
              "state_order": list(self.state_order), **self.hyperparams}
        if not self.spec.rbf:
            hp["activation"] = self.spec.activation
        hp["final_activation"] = self.spec.final_activation
        return hp

    def _own_leaves(self):
        return []

    def save_to_eqx(self, path, name: str, **kwargs):
        """Write the reference's checkpoint format (neuralbase.py:530-571) so that
        ``catalax.neural.<Class>.from_eqx`` (and ours) can load the model."""
        import json
        import os

        if name.endswith(".eqx"):
            name = name[: -len(".eqx")]
        filename = os.path.join(str(path), name + ".eqx")
        hp = self._header()
        f32 = lambda a: np.asarray(a, dtype=np.float32)
        leaves = []
        for l, (W, b) in enumerate(zip(self.weights, self.biases)):
            leaves.append(f32(W))
            if b is not None:
                leaves.append(f32(b))
            if self.spec.rbf and l < self.depth:
                leaves += [f32(self.rbf_mu[l]), np.asarray(self.rbf_gamma, dtype=np.float32)]
        leaves.append(np.asarray(self.max_time, dtype=np.float32))
        leaves += [np.asarray(int(i)) for i in self.observable_indices]

        def numeric_leaves(o):      # jax flattens dicts with sorted keys; str / None are no leaves
            if isinstance(o, dict):
                return [x for key in sorted(o) for x in numeric_leaves(o[key])]
            if isinstance(o, (list, tuple)):
                return [x for v in o for x in numeric_leaves(v)]
            if isinstance(o, (bool, int, float, np.generic)):
                return [np.asarray(o)]
            return []
        leaves += numeric_leaves(hp)
        leaves += self._own_leaves()
        with open(filename, "wb") as f:
            f.write((json.dumps({"hyperparameters": hp, **kwargs}, default=str) + "\n").encode())
            for a in leaves:
                np.save(f, a, allow_pickle=False)
        return filename

    @classmethod
    def from_model(cls, model, width_size: int, depth: int, seed: int = 0,
                   use_final_bias: bool = False, final_activation="identity", solver=Tsit5,
                   activation="tanh", **kwargs):
        """neuralbase.py:392-431: sizes and state order come from a ``Model``."""
        modeled = len(model.odes) > 0 or len(model.reactions) > 0
        state_order = model.get_state_order(modeled=modeled)
        return cls(data_size=len(state_order), width_size=width_size, depth=depth,
                   state_order=state_order,
                   observable_indices=model.get_state_order(as_indices=True, modeled=modeled),
                   solver=solver, activation=activation, final_activation=final_activation,
                   use_final_bias=use_final_bias, key=seed, **kwargs)

    def get_weights_and_biases(self):
        """Every array leaf of the model (neuralbase.py:576-584 flattens the whole pytree): MLP
        weights and biases, RBF centres, ``max_time`` once training made it an array, and the
        class-specific leaves (stoichiometric matrix / mass constraint; mechanistic parameters,
        residual scale, gate matrix)."""
        out = [a for pair in zip(self.weights, self.biases) for a in pair if a is not None]
        if getattr(self, "rbf_mu", None) is not None:
            out += list(self.rbf_mu)
        if getattr(self, "_max_time_is_leaf", False):
            out.append(np.asarray(self.max_time, dtype=np.float64))
        if self.kind == "rateflow":
            out.append(self.stoich_matrix)
            if self.mass_constraint is not None:
                out.append(self.mass_constraint)
        if self.kind == "universalode":
            out += [self.parameters, self.alpha_residual, self.gate_matrix]
        return out

    def n_parameters(self) -> int:
        return int(sum(a.size for pair in zip(self.weights, self.biases) for a in pair if a is not None))

    # ------------------------------------------------------------------ forward
    def predict_arrays(self, ts, y0s, solver=None, rtol: Optional[float] = None,
                       atol: Optional[float] = None, dt0: Optional[float] = None, kernel: int = 0):
        """vmap(lambda ts, y0: self(ts, y0, ...)) (neuralbase.py:307-310): ts [B,T]|[T], y0s [B,S]
        -> ys [B,T,S].  numpy in -> numpy out; torch CUDA in -> torch CUDA out."""
        import torch

        dtype = default_dtype()
        model, theta = self._model(dtype)
        on_device = isinstance(y0s, torch.Tensor) and y0s.is_cuda
        dev = y0s.device if on_device else torch.device("cuda", torch.cuda.current_device())
        tt = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=dtype))
        ts_t, y0_t = tt(ts).to(dev), tt(y0s).to(dev)
        if y0_t.dim() == 1:
            y0_t = y0_t[None]
        if dt0 is None:
            # `dt0 or ts[1] - ts[0]` under vmap (neuralode.py:44-47): every trajectory takes the
            # spacing of ITS time row -- dt0 = 0 asks the kernel for exactly that
            dt0 = 0.0
        sname = solver_name(solver or self.solver)
        if kernel == 0:     # auto: tensor cores (kernel 3) when eligible, else CUDA cores (2)
            kernel = 3 if (self.tensor_core_eligible(dtype) and sname in ("tsit5", "dopri5")) else 2
        if kernel == 3:
            theta = self._tc_theta()
        out = model.solve(y0_t, torch.as_tensor(theta, device=dev),
                          torch.empty((y0_t.shape[0], 0), dtype=y0_t.dtype, device=dev), ts_t,
                          solver=sname,
                          in_axes=(0, None, 0, 0 if ts_t.dim() == 2 else None),
                          rtol=1e-3 if rtol is None else rtol, atol=1e-6 if atol is None else atol,
                          dt0=dt0, max_steps=self._max_steps(), t0_from_ts=self._t0_from_ts(),
                          kernel=kernel)
        self.last = out
        return out.ys if on_device else out.ys.cpu().numpy()

    def __call__(self, ts, y0, solver=None, rtol=None, atol=None, dt0=None):
        """One trajectory: ts [T], y0 [S] -> [T,S] (the per-class __call__ of the reference)."""
        return self.predict_arrays(ts, np.asarray(y0)[None] if not hasattr(y0, "is_cuda") else y0[None],
                                   solver, rtol, atol, dt0)[0]

    def predict(self, dataset, config=None, n_steps: int = 100, use_times: bool = False):
        """Dataset in, Dataset out (neuralbase.py:252-317); times come from the dataset."""
        from ..dataset import Dataset

        y0s = dataset.to_y0_matrix(self.state_order)
        if use_times or config is None:
            times = np.stack([m.time for m in dataset.measurements])
        else:
            times = np.linspace(np.asarray(config.t0, dtype=np.float64),
                                np.asarray(config.t1, dtype=np.float64), config.nsteps or n_steps)
            times = np.broadcast_to(np.asarray(times).T, (len(dataset), times.shape[0])) \
                if np.ndim(times) == 1 else np.asarray(times).T
        ys = self.predict_arrays(times.astype(default_dtype()), y0s)
        return Dataset.from_jax_arrays(self.state_order, ys, times, y0s)

    def rates(self, t, y, constants=None):
        """vmap(field)(t, y) (neuralbase.py:319-327): t [N], y [N,S] -> [N,S]."""
        import torch

        dtype = default_dtype()
        model, theta = self._model(dtype)
        on_device = isinstance(y, torch.Tensor) and y.is_cuda
        dev = y.device if on_device else torch.device("cuda", torch.cuda.current_device())
        tt = lambda a: (a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=dtype))).to(dev)
        y_t = tt(y)
        out = model.rates(tt(t), y_t, torch.as_tensor(theta, device=dev),
                          torch.empty((0,), dtype=y_t.dtype, device=dev))
        return out if on_device else out.cpu().numpy()


    # ---- Surrogate protocol (catalax/surrogate.py:13-73), used by BayesianModel(surrogate=...)
    has_uncertainty = False            # neuralbase.py:146-152 (ensembles only)

    def predict_rates(self, dataset):
        """Rates at every measured point of a Dataset, [M*T, S] (neuralbase.py:329-344)."""
        data, times, _ = dataset.to_jax_arrays(self.state_order)
        M, T, S = data.shape
        return np.asarray(self.rates(np.asarray(times).ravel(), np.asarray(data).reshape(M * T, S)))

    def rate_jacobian(self, t, y, rel_step: float = 1e-6):
        """``J[p, i, k] = d rate_i / d y_k`` at N points, [N,S,S] float64.  The reference takes
        ``jax.jacfwd`` of ``rates`` (mcmc.py:865-894); here: central differences of the SAME
        field evaluated in float64 by the generated kernel (ctx_rates), accurate to ~1e-9."""
        import torch

        model, theta = self._model(np.float64)
        dev = torch.device("cuda", torch.cuda.current_device())
        y = torch.as_tensor(np.asarray(y, dtype=np.float64), device=dev)
        t = torch.as_tensor(np.asarray(t, dtype=np.float64), device=dev)
        th = torch.as_tensor(np.asarray(theta, dtype=np.float64), device=dev)
        c = torch.empty((0,), dtype=torch.float64, device=dev)
        N, S = y.shape
        J = torch.empty((N, S, S), dtype=torch.float64, device=dev)
        for k in range(S):
            h = rel_step * torch.clamp(y[:, k].abs(), min=1.0)
            yp, ym = y.clone(), y.clone()
            yp[:, k] += h
            ym[:, k] -= h
            J[:, :, k] = (model.rates(t, yp, th, c) - model.rates(t, ym, th, c)) / (2.0 * h)[:, None]
        return J.cpu().numpy()


    # ---- training: value and gradient of the loss (trainer.py:268-306)
    def value_and_grad(self, ts, y0, data, *, loss="l2", mask=None, rtol=None, atol=None, dt0=None,
                       want_y0: bool = False, cap: int = 512, solver=None):
        """Loss of ``vmap(self)(ts, y0)`` against ``data`` and its gradient w.r.t. every weight.

        ts [B,T] (or [T]), y0 [B,S], data [B,T,S]; ``mask`` [T] weights the time points (temporal
        dropout, trainer.py:300-306): ``loss = sum(mask * per_point) / (sum(mask) * B * S)``.
        ``loss``: "l2" (``optax.l2_loss``, 0.5 (p - y)^2), "l1", or a callable
        ``(pred, target) -> per-point loss`` written with torch operators.
        Returns ``(value, grads)`` with ``grads = {"weights": [...], "biases": [...]}`` in the
        equinox convention (W [out, in]; None where the model has no bias)."""
        import torch

        dtype = default_dtype()
        model, theta = self._model(dtype)
        if not model.field.source.count("#define CTX_MLP_ADJ 1"):
            raise NotImplementedError("weight gradients: neural fields of <= 8 states whose network "
                                      "fits shared memory")
        sname = solver_name(solver or self.solver)
        if sname not in ("tsit5", "dopri5"):
            raise NotImplementedError("weight gradients are implemented for Tsit5 and Dopri5")
        dev = y0.device if isinstance(y0, torch.Tensor) and y0.is_cuda else \
            torch.device("cuda", torch.cuda.current_device())
        tt = lambda a: (a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=dtype))).to(dev)
        ts_t, y0_t, data_t = tt(ts), tt(y0), tt(data)
        if dt0 is None:
            dt0 = 0.0        # kernel: ts[1] - ts[0] of every trajectory (neuralode.py:44-47)
        ctx = model.adjoint_forward(y0_t, torch.as_tensor(theta, device=dev), ts_t,
                                    rtol=1e-3 if rtol is None else rtol,
                                    atol=1e-6 if atol is None else atol, dt0=dt0,
                                    max_steps=self._max_steps(), cap=cap, solver=sname)
        self.last = ctx
        bad = int((ctx["status"] != 0).sum().item())
        if bad:
            raise RuntimeError(f"{bad} trajectories failed in the training solve "
                               f"(status {int(ctx['status'].max().item())}; 3 = more than cap={cap} steps)")
        pred = ctx["ys"].detach().clone().requires_grad_(True)
        value = masked_loss(pred, data_t.to(pred.dtype), loss, None if mask is None else tt(mask))
        (gys,) = torch.autograd.grad(value, pred)
        g = model.adjoint_backward(ctx, gys, want_y0=want_y0)
        g, gy0 = g if want_y0 else (g, None)
        g = g.to(torch.float64).cpu().numpy()
        lay = cm.layout_of(self.spec)
        gW, gb = [], []
        for l, (n_in, n_out) in enumerate(self.spec.layer_dims()):
            ip = lay.in_pad[l]
            gW.append(g[lay.w_off[l]: lay.w_off[l] + n_out * ip].reshape(n_out, ip)[:, :n_in].copy())
            gb.append(None if self.biases[l] is None else g[lay.b_off[l]: lay.b_off[l] + n_out].copy())
        grads = {"weights": gW, "biases": gb,
                 # the kernel differentiates w.r.t. 1 / max_time (the value its parameter pack carries)
                 "max_time": float(-g[lay.inv_max_time] / (self.max_time * self.max_time))}
        if self.kind == "rateflow":
            S, R_ = self.data_size, self.spec.out_size
            grads["stoich"] = g[lay.stoich: lay.stoich + S * R_].reshape(S, R_).copy()
        if self.kind == "universalode":
            S = self.data_size
            grads["gate_matrix"] = g[lay.gate: lay.gate + S * S].reshape(S, S).copy()
            # the kernel differentiates w.r.t. softplus(alpha_residual); softplus' = sigmoid
            grads["alpha_residual"] = g[lay.alpha: lay.alpha + S] / (1.0 + np.exp(-self.alpha_residual))
            grads["parameters"] = g[lay.mech: lay.mech + len(self.parameters)].copy()
        if want_y0:
            grads["y0"] = gy0
        return float(value.item()), grads

    def apply_updates(self, dW, db, dstoich=None, dgate=None, dalpha=None, dparameters=None,
                      dmax_time=None):
        """weights += dW, biases += db (and the class-specific leaves); the packs are rebuilt."""
        if dmax_time is not None:
            self.max_time = float(self.max_time + dmax_time)
            self._compiled.pop("tc", None)
        if dstoich is not None:
            self.stoich_matrix = self.stoich_matrix + np.asarray(dstoich, dtype=np.float64)
        if dgate is not None:
            self.gate_matrix = self.gate_matrix + np.asarray(dgate, dtype=np.float64)
        if dalpha is not None:
            self.alpha_residual = self.alpha_residual + np.asarray(dalpha, dtype=np.float64)
        if dparameters is not None:
            self.parameters = self.parameters + np.asarray(dparameters, dtype=np.float64)
        self.weights = [w + np.asarray(d, dtype=np.float64) for w, d in zip(self.weights, dW)]
        self.biases = [None if b is None else b + np.asarray(d, dtype=np.float64)
                       for b, d in zip(self.biases, db)]
        for key in list(self._compiled):
            if key == "tc":
                del self._compiled[key]
            else:
                model, _ = self._compiled[key]
                self._compiled[key] = (model, cm.pack_parameters(
                    self.spec, self.weights, self.biases, self.max_time, np.dtype(key),
                    **self._pack_extra()))

    def train(self, ts, y0=None, data=None, *, steps: int = 100, lr: float = 1e-3, loss="l2", **kw):
        """Two call forms.  ``train(dataset, strategy, ...)`` is the reference's
        ``NeuralBase.train`` (neuralbase.py:185-250): the curriculum of ``neural/trainer.py``.
        ``train(ts, y0, data, steps=, lr=)`` is a plain Adam loop over ``value_and_grad`` on arrays
        and returns the loss history."""
        if hasattr(ts, "measurements"):
            from .trainer import train_neural_ode

            strategy = y0 if y0 is not None else kw.pop("strategy")
            return train_neural_ode(self, ts, strategy, **kw)
        m = [[np.zeros_like(w) for w in self.weights], [None if b is None else np.zeros_like(b) for b in self.biases]]
        v = [[np.zeros_like(w) for w in self.weights], [None if b is None else np.zeros_like(b) for b in self.biases]]
        b1, b2, eps, hist = 0.9, 0.999, 1e-8, []
        for k in range(1, steps + 1):
            val, g = self.value_and_grad(ts, y0, data, loss=loss, **kw)
            hist.append(val)
            upd = [[], []]
            for grp, (params, grads) in enumerate(((self.weights, g["weights"]), (self.biases, g["biases"]))):
                for i, (p_, g_) in enumerate(zip(params, grads)):
                    if p_ is None:
                        upd[grp].append(None)
                        continue
                    m[grp][i] = b1 * m[grp][i] + (1 - b1) * g_
                    v[grp][i] = b2 * v[grp][i] + (1 - b2) * g_ * g_
                    mh, vh = m[grp][i] / (1 - b1 ** k), v[grp][i] / (1 - b2 ** k)
                    upd[grp].append(-lr * mh / (np.sqrt(vh) + eps))
            self.apply_updates(upd[0], upd[1])
        return hist


class NeuralODE(NeuralBase):
    """Plain MLP vector field (neuralode.py:25-59)."""

    kind = "neuralode"


class RateFlowODE(NeuralBase):
    """``stoich_matrix @ relu(MLP)`` (rateflow.py:29-121)."""

    kind = "rateflow"

    def __init__(self, data_size: int, reaction_size: int, width_size: int, depth: int,
                 state_order: List[str], observable_indices: Optional[List[int]] = None,
                 solver=Tsit5, activation="celu", use_final_bias: bool = False,
                 learn_stoich: bool = True, stoich_matrix=None, mass_constraint=None, *, key=0,
                 **kw):
        super().__init__(data_size=data_size, width_size=width_size, depth=depth,
                         state_order=state_order, observable_indices=observable_indices,
                         solver=solver, activation=activation, final_activation="relu",
                         use_final_bias=use_final_bias, out_size=reaction_size, key=key,
                         reaction_size=reaction_size, learn_stoich=learn_stoich, **kw)
        self.reaction_size = reaction_size
        self.learn_stoich = learn_stoich
        if mass_constraint is not None:     # rateflow.py:62-75
            mass_constraint = np.asarray(mass_constraint, dtype=np.float64)
            assert mass_constraint.ndim == 2 and mass_constraint.shape[1] == len(state_order), (
                "Mass constraint must be a matrix of shape (n_constraints, n_state). Given shape: "
                f"{mass_constraint.shape}")
        self.mass_constraint = mass_constraint
        self.stoich_matrix = (self._rng.normal(size=(data_size, reaction_size))
                              if stoich_matrix is None else np.asarray(stoich_matrix, dtype=np.float64))

    default_activation = "celu"      # rateflow.py:38

    def _extra(self):
        return {"stoich_matrix": self.stoich_matrix}

    def _own_leaves(self):           # rateflow.py:19-27: stoich_matrix, reaction_size, mass_constraint
        out = [np.asarray(self.stoich_matrix, dtype=np.float32), np.asarray(int(self.reaction_size))]
        if self.mass_constraint is not None:
            out.append(np.asarray(self.mass_constraint, dtype=np.float32))
        return out

    @classmethod
    def from_model(cls, model, width_size: int, depth: int, reaction_size: int, seed: int = 0,
                   use_final_bias: bool = False, solver=Tsit5, activation="celu", **kwargs):
        state_order = model.get_state_order(modeled=len(model.odes) > 0 or len(model.reactions) > 0)
        return cls(data_size=len(state_order), reaction_size=reaction_size, width_size=width_size,
                   depth=depth, state_order=state_order,
                   observable_indices=list(range(len(state_order))), solver=solver,
                   activation=activation, use_final_bias=use_final_bias, key=seed, **kwargs)


class UniversalODE(NeuralBase):
    """``mech + softplus(alpha) * sigmoid(10 G y) * MLP`` (universalode.py:28-141)."""

    kind = "universalode"

    def __init__(self, data_size: int, width_size: int, depth: int, model,
                 observable_indices: Optional[List[int]] = None, solver=Tsit5,
                 use_final_bias: bool = False, activation="celu", final_activation="identity",
                 use_gate: bool = True, *, key=0, gate_matrix=None, alpha_residual=None,
                 parameters=None, **kw):
        if isinstance(model, str):    # universalode.py:47-50
            import json

            model = json.loads(model)
        if isinstance(model, dict):
            from ..model import Model

            model = Model.from_dict(model)
        resolved = model._replace_assignments()
        state_order = model.get_state_order()
        super().__init__(data_size=data_size, width_size=width_size, depth=depth,
                         state_order=state_order, observable_indices=observable_indices,
                         solver=solver, activation=activation, final_activation=final_activation,
                         use_final_bias=use_final_bias, key=key,
                         mech=resolved.sim_input.field_spec(), model=model.to_dict(), **kw)
        self.use_gate = True          # effectively always True in the reference (universalode.py:65-68)
        self.model = model
        self.parameters = np.asarray(resolved._get_parameters() if parameters is None else parameters,
                                     dtype=np.float64)
        S = len(state_order)
        self.alpha_residual = (0.05 * self._rng.normal(size=(S,)) if alpha_residual is None
                               else np.asarray(alpha_residual, dtype=np.float64))
        self.gate_matrix = (0.01 * self._rng.normal(size=(S, S)) if gate_matrix is None
                            else np.asarray(gate_matrix, dtype=np.float64))

    def _extra(self):
        return {"gate_matrix": self.gate_matrix, "alpha_residual": self.alpha_residual,
                "mech_parameters": self.parameters}

    default_activation = "celu"      # universalode.py:38

    def _own_leaves(self):           # universalode.py:20-26: parameters, alpha_residual, gate_matrix
        f32 = lambda a: np.asarray(a, dtype=np.float32)
        return [f32(self.parameters), f32(self.alpha_residual), f32(self.gate_matrix)]

    @classmethod
    def from_model(cls, model, width_size: int, depth: int, seed: int = 0,
                   use_final_bias: bool = False, final_activation="identity", solver=Tsit5,
                   activation="celu", **kwargs):
        """universalode.py:188-: sizes come from the mechanistic ``Model``."""
        n = len(model.get_state_order())
        return cls(data_size=n, width_size=width_size, depth=depth, model=model,
                   observable_indices=list(range(n)), solver=solver, activation=activation,
                   final_activation=final_activation, use_final_bias=use_final_bias, key=seed,
                   **kwargs)

    def _t0_from_ts(self) -> bool:
        return False    # universalode.py:133: t0 = 0.0

    def _max_steps(self) -> int:
        return 64 ** 4  # universalode.py:140
