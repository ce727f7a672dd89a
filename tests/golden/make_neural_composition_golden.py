"""Golden vectors for the neural vector-field COMPOSITIONS, produced by executing the reference's
own source lines (run in the build container, where /root/reference exists; the fixture travels).

jax / equinox cannot be imported here, so the reference classes cannot be instantiated.  What CAN
run is the text of the methods that define the compositions: this script extracts, with `ast`,

  * catalax/neural/mlp.py          MLP.__call__              (t / max_time, concat, self.mlp)
  * catalax/neural/rbf.py          RBFLayer.__call__         (pairwise distances to the centres)
  * catalax/neural/rateflow.py     RateFlowODE.stoich_func   (stoich_matrix @ relu(func))
  * catalax/neural/universalode.py UniversalODE._corrective_term / _combined_term
                                   (mech + softplus(alpha) * sigmoid(10 G y) * func)

and executes them UNMODIFIED against a numpy stand-in for the few jax functions they call
(`jnp.concatenate/array/expand_dims/sum/square/exp`, `jax.nn.relu/sigmoid/softplus`).  Only the
inner `eqx.nn.MLP` (`W @ x + b`, activation after hidden layers, final_activation after the last)
is a restatement -- that convention is pinned separately on the trained checkpoint the reference
ships (tests/test_eqx_checkpoint.py).  Output: tests/golden/neural_composition.npz.
"""
import ast
import json
import sys
import textwrap
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/catalax/neural")
OUT = Path(__file__).parent / "neural_composition.npz"


def method_source(path: Path, cls: str, name: str) -> str:
    tree = ast.parse(path.read_text())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == name:
                    lines = path.read_text().splitlines()[item.lineno - 1: item.end_lineno]
                    src = textwrap.dedent("\n".join(lines))
                    # drop return annotations such as `-> jax.Array` (evaluated at def time)
                    fn = ast.parse(src).body[0]
                    fn.returns = None
                    for a in fn.args.args:
                        a.annotation = None
                    return ast.unparse(fn)
    raise KeyError((path, cls, name))


# ---- numpy stand-ins for the jax names the extracted lines use
jnp = types.SimpleNamespace(concatenate=np.concatenate, array=np.array, expand_dims=np.expand_dims,
                            sum=np.sum, square=np.square, exp=np.exp)
nn = types.SimpleNamespace(relu=lambda x: np.maximum(x, 0.0),
                           sigmoid=lambda x: 1.0 / (1.0 + np.exp(-x)),
                           softplus=lambda x: np.logaddexp(x, 0.0))
jax = types.SimpleNamespace(nn=nn, numpy=jnp, Array=np.ndarray)
NS = {"jax": jax, "jnp": jnp, "jnn": nn, "np": np}


def define(path, cls, name):
    ns = dict(NS)
    exec(method_source(path, cls, name), ns)
    return ns[name]


mlp_call = define(REF / "mlp.py", "MLP", "__call__")
rbf_call = define(REF / "rbf.py", "RBFLayer", "__call__")
stoich_func = define(REF / "rateflow.py", "RateFlowODE", "stoich_func")
corrective = define(REF / "universalode.py", "UniversalODE", "_corrective_term")
combined = define(REF / "universalode.py", "UniversalODE", "_combined_term")

ACT = {"tanh": np.tanh, "celu": lambda x: np.maximum(x, 0.0) + np.expm1(np.minimum(x, 0.0)),
       "softplus": nn.softplus, "identity": lambda x: x, "relu": nn.relu}


class EqxMLP:
    """eqx.nn.MLP semantics (restated): x = layer(x); x = activation(x) for hidden layers,
    final_activation after the last.  With RBF: layers = [Linear, RBF, Linear, RBF, ..., Linear]
    and activation = identity (mlp.py:86-118)."""

    def __init__(self, weights, biases, activation, final_activation, rbf=None):
        self.weights, self.biases = weights, biases
        self.activation, self.final_activation, self.rbf = ACT[activation], ACT[final_activation], rbf

    def __call__(self, x):
        L = len(self.weights)
        for l, (W, b) in enumerate(zip(self.weights, self.biases)):
            x = W @ x + (0.0 if b is None else b)
            if l < L - 1:
                if self.rbf is not None:
                    x = rbf_call(self.rbf[l], x)          # the reference's RBFLayer.__call__ text
                x = self.activation(x)
        return self.final_activation(x)


def make_weights(rng, dims, final_bias):
    Ws = [rng.normal(0, 1 / np.sqrt(i), (o, i)) for i, o in dims]
    bs = [rng.normal(0, 1 / np.sqrt(i), (o,)) for i, o in dims]
    if not final_bias:
        bs[-1] = None
    return Ws, bs


def main():
    rng = np.random.default_rng(20240611)
    S, W, depth, R = 4, 16, 2, 3
    N = 64
    ys = rng.uniform(0.05, 2.0, (N, S))
    ts = rng.uniform(0.0, 10.0, N)
    out = {"ys": ys, "ts": ts, "max_time": np.array(10.0)}

    def func_of(Ws, bs, act, fact, rbf=None):
        obj = types.SimpleNamespace(max_time=10.0, mlp=EqxMLP(Ws, bs, act, fact, rbf))
        return lambda t, y, args: mlp_call(obj, t, y, args)

    # ---- NeuralODE field = MLP.__call__ (tanh) and with RBF hidden layers
    dims = [(S + 1, W), (W, W), (W, S)]
    Ws, bs = make_weights(rng, dims, False)
    f = func_of(Ws, bs, "tanh", "identity")
    out["node_out"] = np.stack([f(t, y, None) for t, y in zip(ts, ys)])
    mus = [rng.uniform(0, 1, (W,)) for _ in range(depth)]
    rbf_layers = [types.SimpleNamespace(mu=m, gamma=0.4) for m in mus]
    f_rbf = func_of(Ws, bs, "identity", "identity", rbf_layers)
    out["rbf_out"] = np.stack([f_rbf(t, y, None) for t, y in zip(ts, ys)])
    for l, (Wm, b) in enumerate(zip(Ws, bs)):
        out[f"node_W{l}"] = Wm
        if b is not None:
            out[f"node_b{l}"] = b
    for l, m in enumerate(mus):
        out[f"rbf_mu{l}"] = m
    out["rbf_gamma"] = np.array(0.4)

    # ---- RateFlowODE: stoich_func (rateflow.py:87-88), MLP 5 -> W -> W -> R, celu, final relu
    dims_r = [(S + 1, W), (W, W), (W, R)]
    Wr, br = make_weights(rng, dims_r, False)
    stoich = rng.normal(size=(S, R))
    rf = types.SimpleNamespace(stoich_matrix=stoich, func=func_of(Wr, br, "celu", "relu"))
    out["rateflow_out"] = np.stack([stoich_func(rf, t, y, None) for t, y in zip(ts, ys)])
    out["rateflow_stoich"] = stoich
    for l, (Wm, b) in enumerate(zip(Wr, br)):
        out[f"rateflow_W{l}"] = Wm
        if b is not None:
            out[f"rateflow_b{l}"] = b

    # ---- UniversalODE: _combined_term (universalode.py:84-101) on the reference's own UODE
    # fixture model (tests/fixtures/uode/uode_model.json), mechanistic field via sympy.lambdify
    import sympy as sp

    md = json.loads(Path("/root/reference/tests/fixtures/uode/uode_model.json").read_text())
    states = sorted(s["symbol"] for s in md.get("states", md.get("species")))
    # inline the assignments, as UniversalODE.__init__ does (universalode.py:71: _replace_assignments)
    loc = {n: sp.Symbol(n) for n in ("I", "E", "oo", "OO", "O", "S", "N")}
    assign = {a["symbol"]: sp.sympify(a["equation"], locals=loc) for a in md.get("assignments", [])}
    for _ in range(len(assign) + 1):
        assign = {k: v.subs({sp.Symbol(q): e for q, e in assign.items() if q != k}) for k, v in assign.items()}
    odes = {o.get("state", o.get("species")): sp.sympify(o["equation"], locals=loc).subs(
        {sp.Symbol(k): v for k, v in assign.items()}) for o in md["odes"]}
    exprs = [odes[s] for s in states]
    free = set().union(*[e.free_symbols for e in exprs])
    params = sorted(str(x) for x in free if str(x) not in states and str(x) != "t")
    pvals = rng.uniform(0.02, 0.2, len(params))
    lam = sp.lambdify([sp.Symbol("t")] + [sp.Symbol(s) for s in states] + [sp.Symbol(p) for p in params],
                      exprs, "numpy")
    vector_field = lambda t, y, args: np.array(lam(t, *y, *args[0]), dtype=np.float64)
    Wu, bu = make_weights(rng, dims, False)
    gate = 0.3 * rng.normal(size=(S, S))
    alpha = 0.5 * rng.normal(size=(S,))
    uo = types.SimpleNamespace(gate_matrix=gate, alpha_residual=alpha, parameters=pvals, use_gate=True,
                               vector_field=vector_field, func=func_of(Wu, bu, "celu", "identity"))
    uo._corrective_term = lambda t, y, args: corrective(uo, t, y, args)
    out["uode_out"] = np.stack([combined(uo, t, y, None) for t, y in zip(ts, ys)])
    out["uode_corrective"] = np.stack([corrective(uo, t, y, None) for t, y in zip(ts, ys)])
    out["uode_gate"], out["uode_alpha"], out["uode_parameters"] = gate, alpha, pvals
    out["uode_param_names"] = np.array(params)
    out["uode_states"] = np.array(states)
    for l, (Wm, b) in enumerate(zip(Wu, bu)):
        out[f"uode_W{l}"] = Wm
        if b is not None:
            out[f"uode_b{l}"] = b
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
