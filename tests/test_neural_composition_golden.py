"""Pin of the neural vector-field COMPOSITIONS on vectors produced by the
reference's own source lines: tests/golden/neural_composition.npz was written by
tests/golden/make_neural_composition_golden.py, which executes -- unmodified, against a numpy
stand-in for jax -- ``MLP.__call__`` (mlp.py:71-84), ``RBFLayer.__call__`` (rbf.py:32-39),
``RateFlowODE.stoich_func`` (rateflow.py:87-88) and ``UniversalODE._corrective_term`` /
``_combined_term`` (universalode.py:76-101).  The oracle (CPU) and the generated CUDA fields (GPU,
through ctx_rates and a short solve) must reproduce them."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).parent / "golden"


def _load():
    g = np.load(GOLD / "neural_composition.npz")
    wb = lambda pre: ([g[f"{pre}_W{l}"] for l in range(3)],
                      [g[f"{pre}_b{l}"] if f"{pre}_b{l}" in g else None for l in range(3)])
    return g, wb


def _uode_mech(g):
    import catalax_b200 as ctx

    mech = ctx.Model.load(GOLD / "reference_fixtures" / "uode_model.json")
    assert mech.get_parameter_order() == list(g["uode_param_names"])
    assert mech.get_state_order() == list(g["uode_states"])
    for name, v in zip(g["uode_param_names"], g["uode_parameters"]):
        mech.parameters[str(name)].value = float(v)
    return mech


def test_oracle_reproduces_the_reference_compositions():
    from oracle import mlp_oracle as mo, symbolic as sy

    g, wb = _load()
    ys, ts = g["ys"], g["ts"]
    W, b = wb("node")
    rhs = mo.make_rhs("neuralode", W, b, "tanh", "identity", 10.0)
    np.testing.assert_allclose(rhs(ts, ys, None, None, None), g["node_out"], rtol=1e-12, atol=1e-13)
    rhs = mo.make_rhs("neuralode", W, b, "identity", "identity", 10.0,
                      rbf_mu=[g["rbf_mu0"], g["rbf_mu1"]], rbf_gamma=float(g["rbf_gamma"]))
    np.testing.assert_allclose(rhs(ts, ys, None, None, None), g["rbf_out"], rtol=1e-12, atol=1e-13)
    W, b = wb("rateflow")
    rhs = mo.make_rhs("rateflow", W, b, "celu", "relu", 10.0, stoich_matrix=g["rateflow_stoich"])
    np.testing.assert_allclose(rhs(ts, ys, None, None, None), g["rateflow_out"], rtol=1e-12, atol=1e-13)
    mech = _uode_mech(g)
    spec = mech._replace_assignments().sim_input.field_spec()
    mech_rhs, *_ = sy.make_rhs(spec.states, spec.parameters, [], spec.odes, spec.reactions)
    W, b = wb("uode")
    rhs = mo.make_rhs("universalode", W, b, "celu", "identity", 10.0, gate_matrix=g["uode_gate"],
                      alpha_residual=g["uode_alpha"], mech_rhs=mech_rhs,
                      mech_parameters=g["uode_parameters"])
    np.testing.assert_allclose(rhs(ts, ys, None, None, ys), g["uode_out"], rtol=1e-11, atol=1e-12)
    # the fixture is not degenerate: the gated correction is a visible part of the field
    assert np.abs(g["uode_corrective"]).max() > 0.05 * np.abs(g["uode_out"]).max()


def _nets(g, wb):
    from catalax_b200.neural import NeuralODE, RateFlowODE, RBFLayer, UniversalODE

    names = ["a", "b", "c", "d"]
    W, b = wb("node")
    node = NeuralODE(4, 16, 2, names, activation="tanh", max_time=10.0, weights=W, biases=b)
    rbf = NeuralODE(4, 16, 2, names, activation=RBFLayer(float(g["rbf_gamma"])), max_time=10.0,
                    weights=W, biases=b, rbf_mu=[g["rbf_mu0"], g["rbf_mu1"]])
    W, b = wb("rateflow")
    rf = RateFlowODE(4, 3, 16, 2, names, activation="celu", max_time=10.0, weights=W, biases=b,
                     stoich_matrix=g["rateflow_stoich"])
    W, b = wb("uode")
    uo = UniversalODE(4, 16, 2, _uode_mech(g), activation="celu", max_time=10.0, weights=W, biases=b,
                      gate_matrix=g["uode_gate"], alpha_residual=g["uode_alpha"])
    return {"node": node, "rbf": rbf, "rateflow": rf, "uode": uo}


@pytest.mark.gpu
def test_cuda_fields_reproduce_the_reference_compositions_f64():
    import catalax_b200 as ctx

    g, wb = _load()
    ctx.enable_x64(True)
    try:
        for key, net in _nets(g, wb).items():
            got = net.rates(g["ts"], g["ys"])
            want = g[f"{key}_out"]
            err = np.abs(got - want).max() / np.abs(want).max()
            assert err < 1e-10, (key, err)
    finally:
        ctx.enable_x64(False)


@pytest.mark.gpu
def test_cuda_fields_reproduce_the_reference_compositions_f32():
    """f32 (the reference's default): 5e-6 of the field's scale for the FFMA form."""
    import catalax_b200 as ctx

    g, wb = _load()
    ctx.enable_x64(False)
    for key, net in _nets(g, wb).items():
        got = net.rates(g["ts"].astype(np.float32), g["ys"].astype(np.float32)).astype(np.float64)
        want = g[f"{key}_out"]
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err < 5e-6, (key, err)


@pytest.mark.gpu
def test_rbf_network_solves_like_the_oracle_f64():
    import catalax_b200 as ctx
    from oracle import diffrax_oracle as do, mlp_oracle as mo

    g, wb = _load()
    ctx.enable_x64(True)
    try:
        net = _nets(g, wb)["rbf"]
        ts = np.linspace(0.5, 10.0, 12)
        ys = net.predict_arrays(ts, g["ys"])
        W, b = wb("node")
        rhs = mo.make_rhs("neuralode", W, b, "identity", "identity", 10.0,
                          rbf_mu=[g["rbf_mu0"], g["rbf_mu1"]], rbf_gamma=float(g["rbf_gamma"]))
        ref = do.solve(rhs, g["ys"], np.zeros(0), np.zeros((64, 0)), ts, rtol=1e-3, atol=1e-6,
                       dt0=float(ts[1] - ts[0]), dtype=np.float64, t0_from_ts=True)
        np.testing.assert_array_equal(net.last.n_accepted.cpu().numpy(), ref.n_accepted)
        np.testing.assert_allclose(ys, ref.ys, rtol=1e-6, atol=1e-9)
    finally:
        ctx.enable_x64(False)
