"""Golden values of the training-loss regularisers, produced by executing the reference's own
``catalax/neural/penalties`` package (run in the build container, where /root/reference exists;
the fixture travels).

The package is imported UNMODIFIED (``__init__``, ``penalties.py``, ``weight.py``, ``uode.py``,
``stoich_mat.py``) with stand-ins in ``sys.modules`` for what it imports: ``jax.numpy`` -> numpy and
``catalax.neural.{neuralbase,rateflow,universalode}`` -> three empty classes (the penalty functions
only ``isinstance``-check them and read attributes).  Values come from the reference functions and
from the reference's ``Penalties.for_*`` presets; gradients are central finite differences of
those same functions (the reference differentiates them with ``jax.grad``).
Output: tests/golden/penalties_reference_lines.npz
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/catalax/neural/penalties")
OUT = Path(__file__).parent / "penalties_reference_lines.npz"

jnp = types.ModuleType("jax.numpy")
for n in ("mean", "where", "abs", "sum", "round", "triu", "ones", "array", "linalg"):
    setattr(jnp, n, getattr(np, n))
jax = types.ModuleType("jax")
jax.numpy = jnp
jax.Array = np.ndarray


class NeuralBase:
    def get_weights_and_biases(self):
        return self._leaves()


class RateFlowODE(NeuralBase):
    pass


class UniversalODE(NeuralBase):
    pass


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


fakes = {"jax": jax, "jax.numpy": jnp,
         "catalax": _module("catalax", __path__=[]), "catalax.neural": _module("catalax.neural", __path__=[]),
         "catalax.neural.neuralbase": _module("catalax.neural.neuralbase", NeuralBase=NeuralBase),
         "catalax.neural.rateflow": _module("catalax.neural.rateflow", RateFlowODE=RateFlowODE),
         "catalax.neural.universalode": _module("catalax.neural.universalode", UniversalODE=UniversalODE)}
saved = {k: sys.modules.get(k) for k in fakes}
sys.modules.update(fakes)
try:
    spec = importlib.util.spec_from_file_location("catalax.neural.penalties", REF / "__init__.py",
                                                  submodule_search_locations=[str(REF)])
    ref = importlib.util.module_from_spec(spec)
    sys.modules["catalax.neural.penalties"] = ref
    spec.loader.exec_module(ref)
    ref_stoich = sys.modules["catalax.neural.penalties.stoich_mat"]
finally:
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def fd_grad(fun, model, attr, h=1e-6):
    """Central finite differences of fun(model) w.r.t. the array attribute `attr`."""
    x = getattr(model, attr)
    g = np.zeros_like(x)
    for idx in np.ndindex(*x.shape):
        keep = x[idx]
        x[idx] = keep + h; hi = float(fun(model))
        x[idx] = keep - h; lo = float(fun(model))
        x[idx] = keep
        g[idx] = (hi - lo) / (2 * h)
    return g


def main():
    rng = np.random.default_rng(7)
    out = {}
    S, R, W = 4, 5, 6
    weights = [rng.normal(0, 0.5, (W, S + 1)), rng.normal(0, 0.5, (W,)), rng.normal(0, 0.5, (R, W))]

    # ---- RateFlowODE: stoichiometric-matrix penalties
    rf = RateFlowODE()
    rf.stoich_matrix = rng.normal(0, 1.0, (S, R))
    rf.mass_constraint = rng.normal(0, 1.0, (2, S))
    rf._leaves = lambda: weights + [rf.stoich_matrix, rf.mass_constraint]
    out["rf_stoich"], out["rf_mass"] = rf.stoich_matrix.copy(), rf.mass_constraint.copy()
    for i, w in enumerate(weights):
        out[f"w{i}"] = w
    funs = {"density": ref.penalize_density, "bipolar": ref.penalize_non_bipolar,
            "conservation": ref.penalize_non_conservative,
            "duplicate_reactions": ref.penalize_duplicate_reactions, "integer": ref.penalize_non_integer,
            "sparsity": ref.l1_stoich_penalty, "null_space": ref_stoich.penalize_null_space,
            "l2": ref.l2_regularisation, "l1": ref.l1_regularisation}
    for name, f in funs.items():
        out[f"rf/{name}/value"] = np.array(float(f(model=rf, alpha=0.37)))
        out[f"rf/{name}/grad_stoich"] = fd_grad(lambda m: f(model=m, alpha=0.37), rf, "stoich_matrix")
    presets = {"default": dict(), "all": dict(alpha=0.2, density_alpha=0.3, bipolar_alpha=0.05,
                                              conservation_alpha=0.4, duplicate_reactions_alpha=1.5,
                                              integer_alpha=0.15, sparsity_alpha=0.01, l2_alpha=1e-3,
                                              null_space_alpha=0.6)}
    for name, kw in presets.items():
        pen = ref.Penalties.for_rateflow(**kw)
        out[f"rf/preset_{name}/names"] = np.array([p.name for p in pen.penalties])
        out[f"rf/preset_{name}/value"] = np.array(float(pen(rf)))
        out[f"rf/preset_{name}/grad_stoich"] = fd_grad(lambda m: pen(m), rf, "stoich_matrix")
    upd = ref.Penalties.for_rateflow().update_alpha(0.05, l2=0.5)
    out["rf/update_alpha/value"] = np.array(float(upd(rf)))
    rf.mass_constraint = None
    rf._leaves = lambda: weights + [rf.stoich_matrix]
    out["rf/null_space_without_constraint"] = np.array(float(ref_stoich.penalize_null_space(model=rf, alpha=0.37)))

    # ---- UniversalODE: gate / residual penalties
    uo = UniversalODE()
    uo.gate_matrix = rng.normal(0, 0.3, (S, S))
    uo.alpha_residual = rng.normal(0, 0.3, (S,))
    uo.parameters = rng.uniform(0.1, 2.0, (3,))
    uo._leaves = lambda: weights + [uo.parameters, uo.alpha_residual, uo.gate_matrix]
    out.update(uo_gate=uo.gate_matrix.copy(), uo_alpha=uo.alpha_residual.copy(), uo_parameters=uo.parameters.copy())
    for name, f in {"l2_gate": ref.l2_reg_gate, "l1_gate": ref.l1_reg_gate, "l2_residual": ref.l2_reg_alpha,
                    "l1_residual": ref.l1_reg_alpha, "l2_mlp": ref.l2_regularisation,
                    "l1_mlp": ref.l1_regularisation}.items():
        out[f"uo/{name}/value"] = np.array(float(f(model=uo, alpha=0.21)))
        out[f"uo/{name}/grad_gate"] = fd_grad(lambda m: f(model=m, alpha=0.21), uo, "gate_matrix")
        out[f"uo/{name}/grad_alpha"] = fd_grad(lambda m: f(model=m, alpha=0.21), uo, "alpha_residual")
    for name, kw in {"default": dict(), "l1": dict(l1_gate_alpha=0.02, l1_residual_alpha=0.03, l1_mlp_alpha=0.004,
                                                     l2_gate_alpha=None)}.items():
        pen = ref.Penalties.for_universal_ode(**kw)
        out[f"uo/preset_{name}/names"] = np.array([p.name for p in pen.penalties])
        out[f"uo/preset_{name}/value"] = np.array(float(pen(uo)))

    # ---- NeuralODE preset
    no = NeuralBase()
    no._leaves = lambda: weights
    for name, kw in {"default": dict(), "both": dict(l2_alpha=2e-3, l1_alpha=5e-4)}.items():
        pen = ref.Penalties.for_neural_ode(**kw)
        out[f"no/preset_{name}/names"] = np.array([p.name for p in pen.penalties])
        out[f"no/preset_{name}/value"] = np.array(float(pen(no)))
        out[f"no/preset_{name}/grad_w0"] = fd_grad(
            lambda m: pen(m), types.SimpleNamespace(w=weights[0], get_weights_and_biases=lambda: weights), "w")
    # ---- the training loss: trainer.py:257-311 (_prepare_step_and_loss -> grad_loss) executed
    # unmodified; eqx.filter_value_and_grad / filter_jit -> pass-through decorators, eqx.combine ->
    # the differentiable part, jax.vmap(model, in_axes=(0, 0)) -> a loop, optax.l2_loss -> its
    # documented formula 0.5 (p - y)^2, the penalties -> the reference's own for_neural_ode preset
    import ast
    ttext = Path("/root/reference/catalax/neural/trainer.py").read_text()
    tlines = ttext.splitlines()
    fn = next(n for n in ast.parse(ttext).body if isinstance(n, ast.FunctionDef) and n.name == "_prepare_step_and_loss")
    eqx = types.SimpleNamespace(filter_value_and_grad=lambda f: f, filter_jit=lambda f: f,
                                combine=lambda diff, static: diff, filter=None, is_inexact_array=None,
                                apply_updates=None)
    jax_t = types.SimpleNamespace(vmap=lambda f, in_axes=None: (lambda a, b: np.stack([f(x, y) for x, y in zip(a, b)])),
                                  Array=np.ndarray)
    ns_t = {"eqx": eqx, "jax": jax_t, "Tuple": tuple, "Callable": object, "Penalties": ref.Penalties}
    exec(compile("from __future__ import annotations\n" + "\n".join(tlines[fn.lineno - 1: fn.end_lineno]),
                 "trainer.py", "exec"), ns_t)
    for loss_name, loss_fun in (("l2", lambda p_, y_: 0.5 * (p_ - y_) ** 2), ("l1", lambda p_, y_: np.abs(p_ - y_))):
        _, grad_loss = ns_t["_prepare_step_and_loss"](loss_fun)
        Bq, Tq, Sq = 5, 7, 3
        pred = rng.normal(size=(Bq, Tq, Sq)); target = rng.normal(size=(Bq, Tq, Sq))
        mask = np.concatenate([np.ones(1), (rng.uniform(size=Tq - 1) < 0.6).astype(float)])
        model_t = types.SimpleNamespace(get_weights_and_biases=lambda: weights)
        call = lambda t_, y0_: pred[int(y0_[0])]
        ti, y0i = np.zeros((Bq, Tq)), np.arange(Bq, dtype=float)[:, None]
        class Diff:
            def __call__(self, t_, y0_): return call(t_, y0_)
            def get_weights_and_biases(self): return weights
        pen = ref.Penalties.for_neural_ode(l2_alpha=1e-3)
        total = grad_loss(Diff(), None, ti, target, y0i, pen, mask)
        out[f"loss/{loss_name}/pred"], out[f"loss/{loss_name}/target"], out[f"loss/{loss_name}/mask"] = pred, target, mask
        out[f"loss/{loss_name}/total"] = np.array(float(total))
        out[f"loss/{loss_name}/penalty"] = np.array(float(pen(Diff())))
    np.savez_compressed(OUT, **out)
    for k in sorted(out):
        if k.endswith("value") or k.endswith("names"):
            print(k, out[k])


if __name__ == "__main__":
    main()
