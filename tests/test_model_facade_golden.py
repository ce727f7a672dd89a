"""Host-side Model rules against records produced by the reference's REAL
``catalax.model.Model`` class (tests/golden/make_model_facade_golden.py imports the reference's
model package unmodified over stub modules for the third-party packages that are not installed):
ordering of states / parameters / constants / reactions, ``{state}_init`` symbols, observable
states, inlining of nested assignments into ODEs and reaction laws, the parameter vector."""
import json
from pathlib import Path

import pytest
import sympy as sp

G = json.loads((Path(__file__).parent / "golden" / "model_facade_reference.json").read_text())


def _build(case):
    import catalax_b200 as ctx

    if case == "assignments":
        m = ctx.Model(name="assignment + constant + init symbol")
        m.add_state(S="Substrate", E="Enzyme", P="Product")
        m.add_constant(I="Inhibitor")
        m.add_assignment("K_app", "K_m * (1 + I / K_i)")
        m.add_assignment("v", "k_cat * E * S / (K_app + S)")
        m.add_ode("S", "-v")
        m.add_ode("P", "v * S_init / S_init")
        m.add_ode("E", "0", observable=False)
        return m, {"K_m": 3.0, "k_cat": 0.5, "K_i": 2.0}
    if case == "init_constant":
        m = ctx.Model(name="init symbol declared as constant")
        m.add_state("s1, s2")
        m.add_constant(s1_init="initial s1", c0="a constant")
        m.add_ode("s1", "-k * s1 * c0")
        m.add_ode("s2", "k * s1_init - q * s2")
        return m, {"k": 0.1, "q": 2.0}
    if case == "reactions":
        m = ctx.Model(name="reactions")
        m.add_state(A="a", B="b", C="c", D="unused")
        m.add_reaction("A + 2 B -> C", symbol="r2", equation="k2 * A * B**2")
        m.add_reaction(symbol="r1", reactants=[(1.0, "C")], products=[(1.0, "A"), (0.5, "B")],
                       equation="k1 * C / (K + C)", reversible=True)
        return m, {"k1": 1.5, "k2": 0.25, "K": 4.0}
    if case == "mixed":
        m = ctx.Model(name="mixed")
        m.add_state(X="x", Y="y", Z="z")
        m.add_assignment("flux", "vmax * X / (km + X)")
        m.add_reaction("X -> Y", symbol="conv", equation="flux")
        m.add_ode("Z", "kz * Y - dz * Z")
        return m, {"vmax": 2.0, "km": 1.0, "kz": 0.3, "dz": 0.7}
    d = json.loads((Path(__file__).parent / "golden" / "reference_fixtures" / "uode_model.json").read_text())
    return ctx.Model.from_dict(d), None


def _parse(text: str):
    """sympify with every identifier a plain Symbol (E, S, I ... are model symbols, not constants)."""
    import re

    names = set(re.findall(r"[A-Za-z_][A-Za-z_0-9]*", text)) - {"exp", "log", "sqrt", "sin", "cos", "tan", "Abs"}
    return sp.sympify(text, locals={n: sp.Symbol(n) for n in names})


def _same_expr(a: str, b) -> bool:
    return sp.simplify(_parse(a) - _parse(str(b))) == 0


@pytest.mark.parametrize("case", [k for k in G if not k.startswith("__")])
def test_model_rules_match_the_reference_class(case):
    m, values = _build(case)
    want = G[case]
    for name, v in (values or {}).items():
        m.parameters[name].value = v
    assert m.get_state_order() == want["state_order"]
    assert m.get_state_order(modeled=False) == want["state_order_unmodeled"]
    assert m.get_state_order(as_indices=True) == want["state_indices"]
    assert m.get_parameter_order() == want["parameter_order"]
    assert m.get_constants_order() == want["constants_order"]
    assert m.get_observable_state_order() == want["observable_states"]
    assert m.get_observable_state_order(as_indices=True) == want["observable_indices"]
    resolved = m._replace_assignments()
    assert sorted(resolved.odes) == sorted(want["resolved_odes"])
    for state, eq in want["resolved_odes"].items():
        assert _same_expr(eq, resolved.odes[state].equation), (state, eq, resolved.odes[state].equation)
    if "reaction_order" in want:
        assert m.get_reaction_order() == want["reaction_order"]
        assert m.get_reacting_state_order() == want["reacting_states"]
        for sym, r in want["reactions"].items():
            mine = resolved.reactions[sym]
            assert [[float(e.stoichiometry), e.state] for e in mine.reactants] == r["reactants"]
            assert [[float(e.stoichiometry), e.state] for e in mine.products] == r["products"]
            assert bool(mine.reversible) == r["reversible"]
            assert _same_expr(r["equation"], mine.equation)
    if "parameter_vector" in want:
        assert [float(x) for x in m._get_parameters()] == pytest.approx(want["parameter_vector"], rel=1e-6)   # (f32 default dtype here)


def _nan(a):
    import numpy as np

    return np.array([[np.nan if v is None else v for v in row] if not isinstance(row[0], list)
                     else [[np.nan if v is None else v for v in r] for r in row] for row in a], dtype=float)


def test_dataset_arrays_match_the_reference_classes():
    """ragged measurements, a state without data, ``pad`` (NaN-padded data, the time
    grid continued), ``to_jax_arrays`` in any state order, ``to_y0_matrix``."""
    import warnings

    import numpy as np

    import catalax_b200 as ctx

    want = G["__dataset__"]
    ds = ctx.Dataset(id="d", name="n", states=["A", "B", "C"])
    ds.add_measurement(ctx.Measurement(id="m1", initial_conditions={"A": 1.0, "B": 0.5, "C": 0.0},
                                       time=np.linspace(0, 4, 5), data={"A": np.arange(5.0), "B": np.arange(5.0) * 2}))
    ds.add_measurement(ctx.Measurement(id="m2", initial_conditions={"A": 2.0, "B": 0.1, "C": 0.3},
                                       time=np.linspace(0, 2, 3), data={"A": np.arange(3.0) + 10, "B": np.arange(3.0) + 20}))
    ds.add_measurement(ctx.Measurement(id="m3", initial_conditions={"A": 3.0, "B": 0.2, "C": 0.4},
                                       time=np.array([0.0, 0.5, 1.5, 4.0]),
                                       data={"A": np.array([1.0, np.nan, 3.0, 4.0]), "B": np.array([5.0, 6.0, 7.0, 8.0])}))
    assert ds.get_observable_states_order() == want["observable_states"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        padded = ds.pad()
        for key, order in (("padded_AB", ["A", "B"]), ("padded_BAC", ["B", "A", "C"])):
            data, time, y0 = padded.to_jax_arrays(order, inits_to_array=True)
            np.testing.assert_allclose(np.asarray(data, dtype=float), _nan(want[key]["data"]), rtol=1e-6, equal_nan=True)
            np.testing.assert_allclose(np.asarray(time, dtype=float), _nan(want[key]["time"]), rtol=1e-6, equal_nan=True)
            np.testing.assert_allclose(np.asarray(y0, dtype=float), _nan(want[key]["y0"]), rtol=1e-6)
        _, _, y0d = padded.to_jax_arrays(["A", "B"])
        assert [dict(d) for d in y0d] == want["inits_dicts"]
    np.testing.assert_allclose(np.asarray(ds.to_y0_matrix(state_order=["A", "B", "C"])), _nan(want["y0_matrix_ABC"]), rtol=1e-6)
    np.testing.assert_allclose(np.asarray(ds.to_y0_matrix(state_order=["C", "A"])), _nan(want["y0_matrix_CA"]), rtol=1e-6)


def test_augment_matches_the_reference_classes():
    """dataset.py:1201-1250 / measurement.py:759-817: augmentation i uses seed + i for every
    measurement, multiplicative noise is ``x * (1 + sigma * n)``, ``append`` keeps the originals in
    front -- with the Threefry normal of PRNGKey(seed) (f32)."""
    import numpy as np

    import catalax_b200 as ctx

    ctx.enable_x64(False)
    ds = ctx.Dataset(id="d", name="n", states=["A", "B"])
    for mid, scale in (("a1", 1.0), ("a2", 3.0)):
        ds.add_measurement(ctx.Measurement(id=mid, initial_conditions={"A": 1.0, "B": 2.0}, time=np.linspace(0, 1, 7),
                                           data={"A": (scale * np.linspace(1, 2, 7)).astype(np.float32),
                                                 "B": (scale * np.linspace(3, 5, 7)).astype(np.float32)}))
    for tag, kw in (("mult_append", dict(n_augmentations=2, sigma=0.05, seed=0, append=True, multiplicative=True)),
                    ("add_only", dict(n_augmentations=1, sigma=0.3, seed=7, append=False, multiplicative=False))):
        want = G["__dataset__"][f"augment_{tag}"]
        aug = ds.augment(**kw)
        assert len(aug.measurements) == want["n"]
        got = np.array([[np.asarray(m.data[s], dtype=np.float64) for s in ("A", "B")] for m in aug.measurements])
        np.testing.assert_allclose(got, np.array(want["data"]), rtol=2e-7, atol=0)
