"""K3/K4: neural vector fields (NeuralODE / RateFlowODE / UniversalODE) vs the oracle.

BASELINE configs[4] at test size: S=4, width 64, depth 3, ts=linspace(0,10,16), max_time=10,
rtol=1e-3, atol=1e-6, dt0=ts[1]-ts[0] (neuralbase.py:634-637), weights rng(3) ~ N(0, 1/fan_in).
x64: 1e-6 relative (+1e-3*atol); f32: stated bound vs the x64 result."""
from pathlib import Path

import numpy as np
import pytest

import catalax_b200 as ctx
from catalax_b200.neural import NeuralODE, RateFlowODE, UniversalODE

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

STATES = ["s0", "s1", "s2", "s3"]
GOLD = Path(__file__).parent / "golden" / "reference_fixtures"


def _make(kind, activation):
    if kind == "neuralode":
        return NeuralODE(4, 64, 3, STATES, activation=activation, max_time=10.0, key=3)
    if kind == "rateflow":
        return RateFlowODE(4, 3, 64, 3, STATES, activation=activation, max_time=10.0, key=3)
    mech = ctx.Model.load(GOLD / "uode_model.json")
    for i, p in enumerate(mech.get_parameter_order()):
        mech.parameters[p].value = 0.05 + 0.01 * i
    return UniversalODE(4, 64, 3, mech, activation=activation, max_time=10.0, key=3)


def _oracle(net, ts, y0s, dtype, t0_from_ts, max_steps):
    from oracle import diffrax_oracle as do, mlp_oracle as mo, symbolic as sy

    kw = {}
    if net.kind == "rateflow":
        kw["stoich_matrix"] = net.stoich_matrix
    if net.kind == "universalode":
        m = net.spec.mech
        mech_rhs, *_ = sy.make_rhs(m.states, m.parameters, [], m.odes, m.reactions)  # p0 is unused
        kw.update(gate_matrix=net.gate_matrix, alpha_residual=net.alpha_residual,
                  mech_rhs=mech_rhs, mech_parameters=net.parameters)
    rhs = mo.make_rhs(net.kind, net.weights, net.biases, net.spec.activation,
                      net.spec.final_activation, net.max_time, **kw)
    B = y0s.shape[0]
    return do.solve(rhs, y0s, np.zeros(0), np.zeros((B, 0)), ts, rtol=1e-3, atol=1e-6,
                    dt0=float(ts[1] - ts[0]), max_steps=max_steps, dtype=dtype,
                    t0_from_ts=t0_from_ts)


@pytest.mark.parametrize("kind,activation", [("neuralode", "tanh"), ("neuralode", "softplus"),
                                             ("rateflow", "celu"), ("universalode", "celu")])
def test_neural_solve_matches_oracle_f64(kind, activation):
    ctx.enable_x64(True)
    try:
        net = _make(kind, activation)
        rng = np.random.default_rng(4)
        B = 200
        y0s = rng.uniform(0.0, 2.0, (B, 4))
        if kind == "universalode":
            y0s = np.concatenate([y0s, ], axis=1)
        ts = np.linspace(0.0 if kind == "universalode" else 0.5, 10.0, 16)
        ys = net.predict_arrays(ts, y0s)
        ref = _oracle(net, ts, y0s, np.float64, net._t0_from_ts(), net._max_steps())
        assert (net.last.status.cpu().numpy() == ref.status).all()
        same = (net.last.n_accepted.cpu().numpy() == ref.n_accepted)
        assert same.mean() > 0.98
        fin = np.isfinite(ref.ys)
        np.testing.assert_array_equal(np.isfinite(ys), fin)
        bound = 1e-6 * np.abs(ref.ys[fin]) + 1e-9
        ok = np.abs(ys[fin] - ref.ys[fin]) <= bound
        # rtol=1e-3: an accept/reject flip from a 1e-16 difference moves a trajectory by ~1e-4
        assert ok.mean() > 0.98
        assert np.abs(ys[fin] - ref.ys[fin]).max() <= 5e-3 * (1 + np.abs(ref.ys[fin]).max())
    finally:
        ctx.enable_x64(False)


def test_neural_rates_match_oracle():
    from oracle import mlp_oracle as mo

    ctx.enable_x64(True)
    try:
        net = _make("rateflow", "celu")
        rng = np.random.default_rng(5)
        N = 1000
        y = rng.uniform(0, 2, (N, 4)); t = rng.uniform(0, 10, N)
        got = net.rates(t, y)
        rhs = mo.make_rhs("rateflow", net.weights, net.biases, "celu", "relu", 10.0,
                          stoich_matrix=net.stoich_matrix)
        want = rhs(t, y, None, None, None)
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
    finally:
        ctx.enable_x64(False)


def test_neural_f32_stated_bound():
    """f32 (the reference default for neural models): |y32 - y64| <= 20*(atol + rtol*|y|)
    for >= 99% of the points, median <= 1x."""
    net = _make("neuralode", "tanh")
    rng = np.random.default_rng(6)
    y0s = rng.uniform(0.0, 2.0, (512, 4)).astype(np.float32)
    ts = np.linspace(0.0, 10.0, 16).astype(np.float32)
    ctx.enable_x64(False)
    y32 = net.predict_arrays(ts, y0s).astype(np.float64)
    ctx.enable_x64(True)
    try:
        y64 = net.predict_arrays(ts.astype(np.float64), y0s.astype(np.float64))
    finally:
        ctx.enable_x64(False)
    scale = 1e-6 + 1e-3 * np.abs(y64)
    err = np.abs(y32 - y64) / scale
    # a random-weight field has unstable directions: a few trajectories amplify the f32/f64
    # difference exponentially over t in [0,10], so the bound is stated on quantiles
    assert (err <= 20).mean() >= 0.99 and np.median(err) <= 1.0


def test_model_rates_kernel_mechanistic():
    """K4 for a mechanistic model: vmap(stack) at N points (model.py:1027-1051)."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle import symbolic as sy
    from tests import cases

    model = cases.nonautonomous()
    rng = np.random.default_rng(1)
    N = 500
    y = rng.uniform(0.2, 2, (N, 4)); th = rng.uniform(0.1, 1, 3); c = rng.uniform(0, 1, (N, 1))
    t = rng.uniform(0, 5, N)
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), "float64")
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    got = m.rates(tt(t), tt(y), tt(th), tt(c), y0=tt(y)).cpu().numpy()
    rhs, *_ = sy.make_rhs(*model[:4], model[4])
    want = rhs(t, y, np.broadcast_to(th, (N, 3)), c, y)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("kind,activation", [("neuralode", "tanh"), ("rateflow", "celu"),
                                             ("universalode", "celu")])
def test_tensor_core_kernel_matches_cuda_core_and_x64(kind, activation):
    """K3 on tcgen05 (kernel 3, 3xTF32 split) vs the CUDA-core form (kernel 2) and vs x64.
    f32 stated bound: |y - y_x64| <= 20*(atol + rtol*|y|) for >= 99% of the points; the two f32
    kernels agree with each other at the same level (median <= 0.2 tolerance units)."""
    net = _make(kind, activation)
    assert net.tensor_core_eligible(np.float32)
    rng = np.random.default_rng(7)
    B = 1000   # ragged: 3 full tiles of 128 + remainder, several CTAs
    y0s = rng.uniform(0.0, 2.0, (B, 4)).astype(np.float32)
    ts = np.linspace(0.0 if kind == "universalode" else 0.5, 10.0, 16).astype(np.float32)
    ctx.enable_x64(False)
    y_cc = net.predict_arrays(ts, y0s, kernel=2).astype(np.float64)
    y_tc = net.predict_arrays(ts, y0s, kernel=3).astype(np.float64)
    st = net.last.status.cpu().numpy()
    ctx.enable_x64(True)
    try:
        y64 = net.predict_arrays(ts.astype(np.float64), y0s.astype(np.float64))
    finally:
        ctx.enable_x64(False)
    assert (st == 0).all() and np.isfinite(y_tc).all()
    np.testing.assert_array_equal(y_tc[:, 0, :] if kind != "universalode" else y_tc[:, 0, :],
                                  y0s.astype(np.float64))       # ts[0] = t0 returns y0 exactly
    scale = 1e-6 + 1e-3 * np.abs(y64)
    e_tc = np.abs(y_tc - y64) / scale
    e_cc = np.abs(y_cc - y64) / scale
    # (random-weight rate fields have unstable directions: a few rows amplify any f32 difference,
    #  for the FFMA kernel just the same -- so the quantile is also compared with that kernel's)
    assert (e_tc <= 20).mean() >= min(0.99, (e_cc <= 20).mean() - 0.01) and np.median(e_tc) <= 1.0
    assert np.median(e_tc) <= 2.0 * np.median(e_cc) + 0.05      # as accurate as the FFMA form
    assert np.median(np.abs(y_tc - y_cc) / scale) <= 0.2


def test_auto_dispatch_uses_tensor_cores_for_f32():
    from catalax_b200 import engine

    net = _make("neuralode", "tanh")
    y0s = np.random.default_rng(0).uniform(0, 2, (300, 4)).astype(np.float32)
    ts = np.linspace(0, 10, 16).astype(np.float32)
    ctx.enable_x64(False)
    a = net.predict_arrays(ts, y0s)             # auto -> kernel 3
    b = net.predict_arrays(ts, y0s, kernel=3)
    np.testing.assert_array_equal(a, b)
