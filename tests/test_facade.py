"""Host-side mirror of the reference's model / dataset interface (no GPU needed).
Covers the ordering rules (model.py:882-999), scheme parser (reaction.py:31-181),
parameter inference (equation.py:48-89), assignment inlining (model.py:788-820), dataset
array export (dataset.py:390-451) and the stoichiometry conventions (stoich.py)."""
import numpy as np
import pytest
import sympy as sp

import catalax_b200 as ctx
from catalax_b200.model.model import Reaction, resolve_assignments


def _mm():
    m = ctx.Model(name="MM")
    m.add_state(S="Substrate", E="Enzyme", P="Product")
    m.add_ode("S", "-k_cat * E * S / (K_m + S)")
    m.add_ode("P", "k_cat * E * S / (K_m + S)")
    m.add_ode("E", "0")
    return m


def test_ordering_is_python_sorted():
    m = _mm()
    assert m.get_state_order() == ["E", "P", "S"]           # 'E' < 'P' < 'S'
    assert m.get_parameter_order() == ["K_m", "k_cat"]       # 'K' < 'k' (ASCII)
    assert m.get_constants_order() == []
    assert m.get_observable_state_order() == ["E", "P", "S"]
    assert m.get_state_order(as_indices=True) == [0, 1, 2]


def test_sympy_shadowed_names_parse_as_symbols():
    m = _mm()
    assert m.odes["S"].equation.free_symbols == {sp.Symbol(n) for n in ("k_cat", "E", "S", "K_m")}


def test_parameter_inference_and_attribute_access():
    m = _mm()
    m.parameters.K_m.value = 0.05
    m.parameters["k_cat"].value = 0.1
    assert m.n_parameters() == 2
    np.testing.assert_allclose(m._get_parameters(), [0.05, 0.1])
    m2 = ctx.Model(name="x")
    m2.add_state("a")
    m2.add_ode("a", "k*a")
    with pytest.raises(ValueError, match="Missing values for parameters: k"):
        m2._get_parameters()
    with pytest.raises(ValueError, match="already exists an ODE"):
        m2.add_ode("a", "k")


def test_constants_and_init_symbols():
    m = ctx.Model(name="c")
    m.add_state(s="Substrate", p="Product")
    m.add_constant(e="Enzyme")
    m.add_constant(s_init="Initial substrate")
    m.add_ode("s", "-k*e*s")
    m.add_ode("p", "s_init")
    assert m.get_constants_order() == ["e"]
    assert "s_init" not in m.parameters and "e" not in m.parameters
    assert "s" in m.odes["p"].inits
    with pytest.raises(ValueError, match="not found in the model for initial value"):
        m.add_ode("q", "zz_init")


def test_scheme_parser():
    r = Reaction.from_scheme("r1", "2 H2 + O2 <=> 2 H2O", "k*H2*O2", True)
    assert [(e.stoichiometry, e.state) for e in r.reactants] == [(2.0, "H2"), (1.0, "O2")]
    assert [(e.stoichiometry, e.state) for e in r.products] == [(2.0, "H2O")]
    for arrow in ("->", "<->", "=>", "<=>", "<==>", "<===>", ">"):
        rr = Reaction.from_scheme("r", f"A {arrow} B", "k*A", False)
        assert rr.reactants[0].state == "A" and rr.products[0].state == "B"
    with pytest.raises(ValueError, match="No supported arrow"):
        Reaction.from_scheme("r", "A B", "k", False)
    with pytest.raises(ValueError, match="at least one reactant"):
        Reaction.from_scheme("r", " -> B", "k", False)
    with pytest.raises(ValueError, match="not found in states"):
        Reaction.from_scheme("r", "A -> Z", "k", False, states=["A", "B"])


def test_reaction_model_field_and_mixed_rule():
    m = ctx.Model(name="r")
    m.add_state("s1, s2")
    m.add_reaction("s1 -> s2", symbol="r1", equation="k1 * s1")
    m.add_reaction("s2 -> s1", symbol="r2", equation="k2 * s2")
    assert m.sim_input.type == "reaction"
    m.add_ode("s1", "k3")
    assert m.sim_input.type == "mixed"
    exprs = m.sim_input.field_spec().total_exprs()
    k1, k2, k3, s1, s2 = sp.symbols("k1 k2 k3 s1 s2")
    assert sp.simplify(exprs[0] - (-k1 * s1 + k2 * s2 + k3)) == 0
    assert sp.simplify(exprs[1] - (k1 * s1 - k2 * s2)) == 0
    with pytest.raises(ValueError, match="Cannot specify both"):
        m.add_reaction("s1 -> s2", symbol="r3", reactants=[(1, "s1")], equation="k")
    with pytest.raises(ValueError, match="not found in model"):
        m.add_reaction(symbol="r4", reactants=[(1, "zz")], products=[(1, "s1")], equation="k")


def test_assignment_inlining():
    out = resolve_assignments({"a": sp.sympify("b + 1"), "b": sp.sympify("2*c"), "c": sp.sympify("k")})
    assert out["a"] == sp.sympify("2*k + 1")
    with pytest.raises(ValueError):
        resolve_assignments({"a": sp.sympify("b"), "b": sp.sympify("a")})
    m = ctx.Model(name="a")
    m.add_state("s, p")
    m.add_assignment("rate", "vmax * s / (km + s)")
    m.add_ode("s", "-rate")
    m.add_ode("p", "rate")
    assert m.get_parameter_order() == ["km", "vmax"]
    inl = m._replace_assignments()
    assert sp.Symbol("rate") not in inl.odes["s"].equation.free_symbols
    assert sp.Symbol("rate") in m.odes["s"].equation.free_symbols  # original untouched


def test_dataset_array_export_and_roundtrip():
    m = _mm()
    ds = ctx.Dataset.from_model(m)
    ds.add_initial(S=100.0, E=10.0, P=0.0)
    ds.add_initial(S=200.0, E=10.0, P=0.0)
    y0 = ds.to_y0_matrix(m.get_state_order())
    np.testing.assert_array_equal(y0, [[10.0, 0.0, 100.0], [10.0, 0.0, 200.0]])
    assert ds.to_y0_matrix(m.get_constants_order()).shape == (2, 0)
    assert y0.dtype == np.float32
    data = np.arange(2 * 5 * 3, dtype=np.float32).reshape(2, 5, 3)
    time = np.tile(np.linspace(0, 1, 5, dtype=np.float32), (2, 1))
    d2 = ctx.Dataset.from_jax_arrays(m.get_state_order(), data, time, y0)
    dd, tt, ii = d2.to_jax_arrays(m.get_state_order(), inits_to_array=True)
    np.testing.assert_array_equal(dd, data)
    np.testing.assert_array_equal(tt, time)
    np.testing.assert_array_equal(ii, y0)
    cfg = d2.to_config(nsteps=7)
    np.testing.assert_array_equal(cfg.t1, [1.0, 1.0])


def test_inaxes_and_config_defaults():
    assert ctx.InAxes.Y0.value == (0, None, 0, None)
    assert ctx.InAxes.Y0 + ctx.InAxes.TIME == (0, None, 0, 0)
    assert ctx.INITS is ctx.InAxes.Y0 and ctx.PARAMETERS is ctx.InAxes.PARAMETERS
    c = ctx.SimulationConfig(t1=10)
    assert (c.nsteps, c.t0, c.dt0, c.rtol, c.atol, c.max_steps, c.throw) == (None, 0, 0.1, 1e-5, 1e-5, 4096, True)
    assert c.solver is ctx.Tsit5


def test_reference_model_json_loads():
    from pathlib import Path

    m = ctx.Model.load(Path(__file__).parent / "golden" / "reference_fixtures" / "uode_model.json")
    assert m.get_state_order() == ["s0", "s1", "s2", "s3"]
    assert len(m.get_parameter_order()) == 10
    spec = m._replace_assignments().sim_input.field_spec()
    from catalax_b200.tools import codegen as cg

    f = cg.generate_field(spec)
    assert f.S == 4 and f.P == 10 and f.C == 1 and f.uses_t


# The reference's tests/unit/model/test_reaction.py (scheme strings -> reactants / products) and
# tests/unit/model/test_stoich.py (matrix layout), case by case against the facade.
SCHEMES = [
    ("A + B -> C", [(1.0, "A"), (1.0, "B")], [(1.0, "C")]),
    ("A + B <-> C", [(1.0, "A"), (1.0, "B")], [(1.0, "C")]),
    ("2 A + 3 B -> 4 C", [(2.0, "A"), (3.0, "B")], [(4.0, "C")]),
    ("1.5 A + 2.5 B -> 3.0 C", [(1.5, "A"), (2.5, "B")], [(3.0, "C")]),
    ("reactant_1 + reactant_2 -> product_1", [(1.0, "reactant_1"), (1.0, "reactant_2")], [(1.0, "product_1")]),
    ("H2 + O2 -> H2O", [(1.0, "H2"), (1.0, "O2")], [(1.0, "H2O")]),
    ("2 H2 + O2 -> 2 H2O", [(2.0, "H2"), (1.0, "O2")], [(2.0, "H2O")]),
    ("A -> B -> C", [(1.0, "A")], [(1.0, "B")]),                    # first arrow splits, rest ignored
    ("  A  +  B  ->  C  ", [(1.0, "A"), (1.0, "B")], [(1.0, "C")]),
    ("A -> B + C + D", [(1.0, "A")], [(1.0, "B"), (1.0, "C"), (1.0, "D")]),
    ("A <===> B", [(1.0, "A")], [(1.0, "B")]),
    ("1 A + 2.5 B + 3 C -> 4.5 D", [(1.0, "A"), (2.5, "B"), (3.0, "C")], [(4.5, "D")]),
]


@pytest.mark.parametrize("schema,reactants,products", SCHEMES)
def test_scheme_cases_of_the_reference(schema, reactants, products):
    r = Reaction.from_scheme(symbol="r", schema=schema, equation="k*A", reversible="<" in schema)
    assert [(e.stoichiometry, e.state) for e in r.reactants] == reactants
    assert [(e.stoichiometry, e.state) for e in r.products] == products
    assert r.reversible is ("<" in schema) and r.symbol == "r"


@pytest.mark.parametrize("schema,message", [
    ("-> A", "Reaction must have at least one reactant and one product"),
    ("A ->", "Reaction must have at least one reactant and one product"),
    ("A + B", "No supported arrow pattern found"),
    ("A + @invalid -> B", "Could not parse term"),
    ("0 A + 1 B -> C", "Stoichiometries must be positive and not zero"),
])
def test_scheme_errors_of_the_reference(schema, message):
    with pytest.raises(ValueError, match=message):
        Reaction.from_scheme(symbol="r", schema=schema, equation="k", reversible=False)


def test_stoichiometry_matrix_layout():
    """test_stoich.py: rows = species in the given order, columns = reactions in sorted-symbol
    order (stoich.py:47-59), overlapping species accumulate, absent species give zero rows."""
    from catalax_b200.model.stoich import derive_stoich_matrix

    mk = lambda sym, lhs, rhs: Reaction.from_scheme(symbol=sym, schema=f"{lhs} -> {rhs}", equation="k",
                                                    reversible=False)
    np.testing.assert_array_equal(derive_stoich_matrix([mk("r1", "A", "B")], ["A", "B"]), [[-1.0], [1.0]])
    two = [mk("r1", "2 A + B", "C"), mk("r2", "C", "D + E")]
    np.testing.assert_array_equal(derive_stoich_matrix(two, ["A", "B", "C", "D", "E"]),
                                  [[-2, 0], [-1, 0], [1, -1], [0, 1], [0, 1]])
    # reaction ordering: columns follow the sorted symbols whatever the list order
    np.testing.assert_array_equal(derive_stoich_matrix(two[::-1], ["A", "B", "C", "D", "E"]),
                                  [[-2, 0], [-1, 0], [1, -1], [0, 1], [0, 1]])
    np.testing.assert_array_equal(derive_stoich_matrix([mk("r1", "A", "B")], ["A", "B", "X"]),
                                  [[-1.0], [1.0], [0.0]])


# The reference's tests/unit/dataset/test_dataset.py (Dataset.pad: the ragged-input path that
# feeds the NaN-masked likelihood), case by case.
def _meas(id, time, inits, data):
    from catalax_b200.dataset.measurement import Measurement

    return Measurement(id=id, time=time, initial_conditions=inits, data=data)


def test_pad_leaves_equal_lengths_alone_and_copies():
    ds = ctx.Dataset(states=["s1", "s2"], id="test_pad")
    ds.add_measurement(_meas("m1", [0, 1, 2], {"s1": 1.0, "s2": 2.0}, {"s1": [1.0, 1.1, 1.2], "s2": [2.0, 2.1, 2.2]}))
    ds.add_measurement(_meas("m2", [0, 1, 2], {"s1": 3.0, "s2": 4.0}, {"s1": [3.0, 3.1, 3.2], "s2": [4.0, 4.1, 4.2]}))
    padded = ds.pad()
    assert padded is not ds and len(padded.measurements) == 2
    for a, b in zip(padded.measurements, ds.measurements):
        assert a is not b and a.id == b.id
        for s in ("s1", "s2"):
            np.testing.assert_allclose(a.data[s], b.data[s])
        np.testing.assert_allclose(a.time, b.time)


def test_pad_missing_state_and_ragged_lengths():
    ds = ctx.Dataset(states=["s1", "s2"], id="test_pad")
    ds.add_measurement(_meas("m1", [0, 1, 2, 3], {"s1": 1.0, "s2": 2.0},
                             {"s1": [1.0, 1.1, 1.2, 1.3], "s2": [2.0, 2.1, 2.2, 2.3]}))
    ds.add_measurement(_meas("m2", [0, 1], {"s1": 3.0, "s2": 4.0}, {"s1": [3.0, 3.1]}))   # no s2 data
    ds.add_measurement(_meas("m3", [0, 1, 2], {"s1": 5.0, "s2": 6.0}, {"s1": [5.0, 5.1, 5.2], "s2": [6.0, 6.1, 6.2]}))
    padded = ds.pad()
    m1, m2, m3 = padded.measurements
    assert not np.isnan(m1.data["s1"]).any() and not np.isnan(m1.data["s2"]).any()
    np.testing.assert_allclose(m2.data["s1"][:2], [3.0, 3.1])
    assert np.isnan(m2.data["s1"][2:]).all() and len(m2.data["s1"]) == 4
    assert np.isnan(m2.data["s2"]).all() and len(m2.data["s2"]) == 4          # added as NaN array
    np.testing.assert_allclose(m2.time, [0, 1, 2, 3])                          # continues with +1
    np.testing.assert_allclose(m3.data["s2"][:3], [6.0, 6.1, 6.2])
    assert np.isnan(m3.data["s2"][3:]).all() and m3.time.tolist() == [0, 1, 2, 3]
    # the original is untouched
    assert "s2" not in ds.measurements[1].data and len(ds.measurements[1].data["s1"]) == 2
    assert len(ds.measurements[2].time) == 3
    # ... and the padded arrays are what the masked likelihood consumes
    data, times, y0s = padded.to_jax_arrays(["s1", "s2"], inits_to_array=True)
    assert data.shape == (3, 4, 2) and times.shape == (3, 4) and y0s.shape == (3, 2)
    assert np.isfinite(data).sum() == 8 + 2 + 6


def test_pad_measurement_with_initial_conditions_only():
    ds = ctx.Dataset(states=["s1", "s2"], id="test_pad")
    ds.add_measurement(_meas("m1", [0, 1, 2], {"s1": 1.0, "s2": 2.0}, {"s1": [1.0, 1.1, 1.2], "s2": [2.0, 2.1, 2.2]}))
    ds.add_measurement(_meas("m2", None, {"s1": 3.0, "s2": 4.0}, {}))
    m1, m2 = ds.pad().measurements
    assert len(m1.data["s1"]) == 3 and len(m1.data["s2"]) == 3
    assert len(m2.data["s1"]) == 3 and np.isnan(m2.data["s1"]).all() and np.isnan(m2.data["s2"]).all()
    assert m2.time is None


def test_pad_needs_initial_conditions_for_every_state():
    ds = ctx.Dataset(states=["s1", "s2"], id="test_pad")
    ds.measurements.append(_meas("m1", [0, 1], {"s1": 1.0}, {"s1": [1.0, 1.1]}))
    with pytest.raises(ValueError, match="missing definition of initial condition"):
        ds.pad()


def test_augment_draws_the_reference_noise():
    """Dataset.augment / Measurement.augment (dataset.py:1201-1250, measurement.py:759-817): the
    noise is jax.random.normal(PRNGKey(seed + i)) -- bitwise the oracle's restatement, which is
    pinned by known answers and by the HMC.ipynb posterior (tests/test_reference_pins.py)."""
    from oracle import jax_prng

    ctx.enable_x64(False)
    ds = ctx.Dataset(states=["s1", "s2"], id="aug")
    t = np.arange(20, dtype=np.float32)
    ds.add_measurement(_meas("m1", t, {"s1": 100.0, "s2": 1.0}, {"s1": 100.0 - 3 * t, "s2": 1.0 + t}))
    ds.add_measurement(_meas("m2", t, {"s1": 50.0, "s2": 2.0}, {"s1": 50.0 - t, "s2": 2.0 + t}))
    mult = ds.augment(n_augmentations=1, seed=0, sigma=5e-2, append=False, multiplicative=True)   # HMC.ipynb cell 4
    assert len(mult.measurements) == 2 and mult.measurements[0].id != "m1"
    z = jax_prng.normal(0, 20)
    for a, b in zip(mult.measurements, ds.measurements):
        assert a.initial_conditions == b.initial_conditions and np.array_equal(a.time, b.time)
        for s in ("s1", "s2"):                      # the same noise vector for every state and series
            np.testing.assert_array_equal(a.data[s], b.data[s] * (z * np.float32(5e-2) + np.float32(1.0)))
    add = ds.augment(n_augmentations=3, sigma=0.01)                          # NeuralODE.ipynb cell 4
    assert len(add.measurements) == 2 + 3 * 2 and add.measurements[0] is ds.measurements[0]
    for i in range(3):
        zi = jax_prng.normal(42 + i, 20)
        np.testing.assert_array_equal(add.measurements[2 + 2 * i + 1].data["s2"],
                                      ds.measurements[1].data["s2"] + zi * np.float32(0.01))
    assert float(jax_prng.normal(0, 1)[0]) == pytest.approx(1.6226422, abs=2e-6)
