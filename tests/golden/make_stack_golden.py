"""Golden right-hand-side values from the reference's REAL stacks and sympy interpreter
(SURVEY 8 a3 / a4), run in the build container.

Through the import hook of make_model_facade_golden.py the reference's ``Model`` builds its own
``ODEStack`` / ``ReactionStack`` / ``MixedStack`` (catalax/tools/stack.py) out of its own
``SymbolicModule`` expression-tree interpreter (catalax/tools/symbolicmodule.py: the op table
:52-98, positional symbol lookup, ``t`` by name) -- all unmodified; the only substitution is
``jax.numpy`` -> numpy, so every node evaluates with the numpy function of the same name.
For each model the stack is evaluated at random points ``(t, y, theta, constants, y0)``.
Output: tests/golden/stack_reference_values.npz (+ the model definitions as JSON strings inside).
"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_model_facade_golden as mk   # noqa: E402  (installs the stubs, imports the reference Model)

OUT = Path(__file__).parent / "stack_reference_values.npz"

OPS = ["Abs(s0 - s1)", "sign(s0 - 1)", "ceiling(s0)", "floor(s1)", "log(s0 + 1)", "exp(-s0)", "sqrt(s0 + s1)",
       "cos(s0)", "acos(s0 / 4)", "sin(s0)", "asin(s0 / 4)", "tan(s0 / 4)", "atan(s0)", "atan2(s0, s1)",
       "cosh(s0 / 3)", "acosh(s0 + 1)", "sinh(s0 / 3)", "asinh(s0)", "tanh(s0)", "atanh(s0 / 4)", "s0 ** p",
       "s0 ** -2", "s0 ** (3 / 2)", "erf(s0)", "Max(s0, s1, p)", "Min(s0, s1)", "2 * s0 / 3",
       "Max(s0 - s1, 0) * p", "s0 * t + c0 * s1_init"]

DEFS = {
    "op_table": {"states": [f"s{i}" for i in range(len(OPS))], "constants": ["c0"],
                 "odes": {f"s{i}": e for i, e in enumerate(OPS)}, "reactions": []},
    "reactions": {"states": ["A", "B", "C", "D"], "constants": [], "odes": {},
                  "reactions": [{"symbol": "r2", "scheme": "A + 2 B -> C", "equation": "k2 * A * B**2"},
                                {"symbol": "r1", "reactants": [[1.0, "C"]], "products": [[1.0, "A"], [0.5, "B"]],
                                 "equation": "k1 * C / (K + C)"},
                                {"symbol": "r3", "scheme": "C -> D", "equation": "k3 * C * exp(-t / 10)"}]},
    "mixed": {"states": ["X", "Y", "Z"], "constants": ["u"], "odes": {"Z": "kz * Y - dz * Z + u * Z_init"},
              "assignments": {"flux": "vmax * X / (km + X)"},
              "reactions": [{"symbol": "conv", "scheme": "X -> Y", "equation": "flux"}]},
}


def build(Model, d):
    m = Model(name="golden")
    m.add_state(", ".join(d["states"]))
    if d["constants"]:
        m.add_constant(", ".join(d["constants"]))
    for sym, eq in d.get("assignments", {}).items():
        m.add_assignment(sym, eq)
    for r in d["reactions"]:
        if "scheme" in r:
            m.add_reaction(r["scheme"], symbol=r["symbol"], equation=r["equation"])
        else:
            m.add_reaction(symbol=r["symbol"], reactants=[tuple(x) for x in r["reactants"]],
                           products=[tuple(x) for x in r["products"]], equation=r["equation"])
    for state, eq in d["odes"].items():
        m.add_ode(state, eq)
    return m


def main():
    rng = np.random.default_rng(31)
    out = {"definitions": np.array(json.dumps(DEFS))}
    for name, d in DEFS.items():
        m = build(mk.Model, d)
        resolved = m._replace_assignments()
        stack = resolved.stack
        S, P, C = len(m.get_state_order()), len(m.get_parameter_order()), len(m.get_constants_order())
        N = 16
        t = rng.uniform(0.0, 5.0, N)
        y = rng.uniform(0.6, 1.9, (N, S))
        th = rng.uniform(0.5, 2.0, (N, P))
        c = rng.uniform(0.1, 1.0, (N, C))
        y0 = rng.uniform(0.6, 1.9, (N, S))
        vals = np.stack([np.asarray(stack(t[i], y[i], (th[i], c[i], y0[i])), dtype=float) for i in range(N)])
        out.update({f"{name}/t": t, f"{name}/y": y, f"{name}/theta": th, f"{name}/consts": c, f"{name}/y0": y0,
                    f"{name}/dy": vals, f"{name}/state_order": np.array(m.get_state_order()),
                    f"{name}/parameter_order": np.array(m.get_parameter_order()),
                    f"{name}/stack": np.array(type(stack).__name__)})
        print(name, type(stack).__name__, vals.shape, np.isfinite(vals).all())
    np.savez_compressed(OUT, **out)


if __name__ == "__main__":
    main()
