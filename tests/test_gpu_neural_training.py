"""NeuralODE training gradient (catalax/neural/trainer.py:268-306 through
neuralode.py:38-59) -- ctx_mlp_adj_forward / ctx_mlp_adj_backward vs the torch-autograd
restatement of the discrete adjoint (oracle/adjoint_oracle.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _problem(S, width, depth, activation, B, T, seed, t1=4.0, use_final_bias=False):
    from catalax_b200.neural import NeuralODE

    rng = np.random.default_rng(seed)
    net = NeuralODE(S, width, depth, [f"s{i}" for i in range(S)], activation=activation,
                    max_time=t1, key=seed, use_final_bias=use_final_bias)
    # damp the random field so trajectories stay O(1) over [0, t1]
    net.weights[-1] = net.weights[-1] * 0.3
    y0 = rng.uniform(0.2, 1.5, (B, S))
    ts = np.tile(np.linspace(0.0, t1, T), (B, 1))
    ts[1:] += rng.uniform(0, 0.05, (B - 1, 1)) * np.arange(T)[None, :] / T   # per-row grids
    data = y0[:, None, :] * np.exp(-0.4 * ts)[:, :, None] + 0.01 * rng.normal(size=(B, T, S))
    return net, ts, y0, data


def _oracle(net, ts, y0, data, mask=None, loss="l2", rtol=1e-3, atol=1e-6):
    from oracle import adjoint_oracle as ao

    d = torch.tensor(data)
    B, T, S = data.shape
    m = torch.ones(T, dtype=torch.float64) if mask is None else torch.tensor(mask, dtype=torch.float64)

    def loss_fn(ys):
        pp = 0.5 * (ys - d) ** 2 if loss == "l2" else (ys - d).abs()
        return (pp * m[None, :, None]).sum() / (m.sum() * B * S)

    return ao.neuralode_value_and_grad(net.weights, net.biases, net.spec.activation,
                                       net.spec.final_activation, net.max_time, ts, y0, loss_fn,
                                       rtol=rtol, atol=atol)


def _check(g, ref_W, ref_b, rtol):
    for l, (a, b) in enumerate(zip(g["weights"], ref_W)):
        scale = np.abs(b).max() + 1e-300
        assert np.abs(a - b).max() <= rtol * scale, (l, float(np.abs(a - b).max() / scale))
    for l, (a, b) in enumerate(zip(g["biases"], ref_b)):
        assert (a is None) == (b is None)
        if a is not None:
            scale = np.abs(b).max() + 1e-300
            assert np.abs(a - b).max() <= rtol * scale, (l, float(np.abs(a - b).max() / scale))


@pytest.mark.parametrize("S,width,depth,activation,B,fb", [
    (2, 32, 2, "tanh", 5, False),
    (4, 64, 3, "softplus", 9, True),      # BASELINE configs[4] shape (5->64->64->64->4)
    (3, 64, 1, "celu", 3, False),
])
def test_weight_gradient_matches_autograd_of_the_discrete_scheme_f64(S, width, depth, activation, B, fb):
    import catalax_b200 as ctx

    ctx.enable_x64(True)
    try:
        net, ts, y0, data = _problem(S, width, depth, activation, B, 8, seed=S + depth, use_final_bias=fb)
        # tight tolerances: many steps per trajectory, with rejections
        val, g = net.value_and_grad(ts, y0, data, want_y0=True, rtol=1e-7, atol=1e-9)
        rv, rys, rW, rb, rgy0, rna, rnr = _oracle(net, ts, y0, data, rtol=1e-7, atol=1e-9)
        assert rna.min() >= 3 and rna.max() >= 4
        assert np.array_equal(net.last["n_accepted"].cpu().numpy(), rna)
        assert np.array_equal(net.last["n_rejected"].cpu().numpy(), rnr)
        np.testing.assert_allclose(net.last["ys"].cpu().numpy(), rys, rtol=1e-9, atol=1e-11)
        assert abs(val - rv) <= 1e-10 * abs(rv)
        _check(g, rW, rb, 1e-8)
        gy0 = g["y0"].cpu().numpy()
        assert np.abs(gy0 - rgy0).max() <= 1e-8 * np.abs(rgy0).max()
    finally:
        ctx.enable_x64(False)


def test_mask_l1_loss_and_reproducibility_f64():
    import catalax_b200 as ctx

    ctx.enable_x64(True)
    try:
        net, ts, y0, data = _problem(2, 32, 2, "tanh", 6, 10, seed=11)
        mask = np.array([1, 0, 1, 1, 0, 1, 1, 1, 0, 1], dtype=np.float64)   # temporal dropout
        val, g = net.value_and_grad(ts, y0, data, loss="l1", mask=mask)
        val2, g2 = net.value_and_grad(ts, y0, data, loss="l1", mask=mask)
        assert val == val2 and all(np.array_equal(a, b) for a, b in zip(g["weights"], g2["weights"]))
        rv, _, rW, rb, *_ = _oracle(net, ts, y0, data, mask=mask, loss="l1")
        assert abs(val - rv) <= 1e-10 * abs(rv)
        _check(g, rW, rb, 1e-8)
    finally:
        ctx.enable_x64(False)


def test_weight_gradient_f32_stated_bound():
    """f32 (the reference's training dtype): 2e-3 of each layer's largest gradient entry."""
    net, ts, y0, data = _problem(4, 64, 3, "softplus", 16, 8, seed=5)
    val, g = net.value_and_grad(ts.astype(np.float32), y0.astype(np.float32), data.astype(np.float32))
    rv, _, rW, rb, *_ = _oracle(net, ts, y0, data)
    assert abs(val - rv) <= 1e-4 * abs(rv)
    _check(g, rW, rb, 2e-3)


def test_adam_steps_on_the_gradient_reduce_the_loss():
    net, ts, y0, data = _problem(2, 32, 2, "tanh", 32, 8, seed=2)
    hist = net.train(ts, y0, data, steps=40, lr=3e-3)
    assert hist[-1] < 0.5 * hist[0], (hist[0], hist[-1])
    # the forward (predict) path sees the updated weights
    pred = net.predict_arrays(ts.astype(np.float32), y0.astype(np.float32))
    mse = float(0.5 * np.mean((np.asarray(pred) - data) ** 2))
    assert abs(mse - hist[-1]) < 0.2 * hist[-1] + 1e-4


def test_rateflow_weight_and_stoichiometry_gradients_f64():
    """RateFlowODE (rateflow.py:87-88): dy = stoich @ relu(MLP); gradients w.r.t. the MLP weights
    AND the (learnable) stoichiometric matrix."""
    import catalax_b200 as ctx
    from catalax_b200.neural import RateFlowODE
    from oracle import adjoint_oracle as ao

    ctx.enable_x64(True)
    try:
        rng = np.random.default_rng(4)
        S, R, B, T, t1 = 4, 3, 7, 8, 4.0
        net = RateFlowODE(S, R, 64, 3, [f"s{i}" for i in range(S)], activation="celu", max_time=t1, key=4)
        net.stoich_matrix = net.stoich_matrix * 0.05      # mild dynamics: no chaotic amplification of rounding
        net.biases[-1] = None if net.biases[-1] is None else net.biases[-1]
        y0 = rng.uniform(0.2, 1.5, (B, S))
        ts = np.tile(np.linspace(0.0, t1, T), (B, 1))
        data = y0[:, None, :] * np.exp(-0.4 * ts)[:, :, None]
        val, g = net.value_and_grad(ts, y0, data, rtol=1e-7, atol=1e-9)
        d = torch.tensor(data)
        loss_fn = lambda ys: (0.5 * (ys - d) ** 2).mean()
        rv, rys, rW, rb, rgy0, rna, rnr, rS = ao.neuralode_value_and_grad(
            net.weights, net.biases, net.spec.activation, net.spec.final_activation, net.max_time,
            ts, y0, loss_fn, rtol=1e-7, atol=1e-9, stoich=net.stoich_matrix)
        # The field has kinks (relu): an error estimate that lands within rounding of the accept
        # threshold can flip one trajectory's accept / reject path between two implementations,
        # which moves that trajectory by O(rtol).  Same path -> rounding-level agreement.
        same_path = np.array_equal(net.last["n_accepted"].cpu().numpy(), rna) and \
            np.array_equal(net.last["n_rejected"].cpu().numpy(), rnr)
        tol_y, tol_g = (5e-8, 1e-6) if same_path else (2e-6, 5e-5)   # (solver rtol: 1e-7; observed 1e-9 / 2e-7)
        np.testing.assert_allclose(net.last["ys"].cpu().numpy(), rys, rtol=tol_y, atol=1e-11)
        assert abs(val - rv) <= 10 * tol_y * abs(rv)
        _check(g, rW, rb, tol_g)
        assert np.abs(g["stoich"] - rS).max() <= tol_g * np.abs(rS).max()
        # one plain gradient step on everything lowers the loss
        lr = 0.05
        net.apply_updates([-lr * w for w in g["weights"]],
                          [None if b is None else -lr * b for b in g["biases"]], -lr * g["stoich"])
        val2, _ = net.value_and_grad(ts, y0, data, rtol=1e-7, atol=1e-9)
        assert val2 < val
    finally:
        ctx.enable_x64(False)


def test_universalode_gradients_f64():
    """UniversalODE (universalode.py:84-141): mech + softplus(alpha) * sigmoid(10 G y) * MLP,
    t0 = 0; gradients w.r.t. the MLP weights, the gate matrix, alpha_residual and the mechanistic
    parameters (all of them inexact-array leaves the reference's optimiser updates)."""
    import catalax_b200 as ctx
    from catalax_b200.neural import UniversalODE
    from oracle import adjoint_oracle as ao

    ctx.enable_x64(True)
    try:
        model = ctx.Model(name="decay2")
        model.add_state(a="a", b="b")
        model.add_ode("a", "-k1 * a / (K + a)")
        model.add_ode("b", "k1 * a / (K + a) - k2 * b + 0.01 * a_init")
        model.parameters.k1.value, model.parameters.k2.value, model.parameters.K.value = 0.8, 0.3, 1.5
        rng = np.random.default_rng(9)
        B, T, t1 = 6, 9, 3.0
        net = UniversalODE(2, 32, 2, model, activation="celu", max_time=t1, key=9,
                           gate_matrix=0.3 * rng.normal(size=(2, 2)),
                           alpha_residual=0.2 * rng.normal(size=2))
        net.weights[-1] = net.weights[-1] * 0.3
        y0 = rng.uniform(0.5, 2.0, (B, 2))
        ts = np.tile(np.linspace(0.2, t1, T), (B, 1))           # first save after t0 = 0
        data = y0[:, None, :] * np.exp(-0.5 * ts)[:, :, None]
        val, g = net.value_and_grad(ts, y0, data, rtol=1e-7, atol=1e-9)
        d = torch.tensor(data)
        loss_fn = lambda ys: (0.5 * (ys - d) ** 2).mean()
        mech = net.spec.mech
        rv, rys, rW, rb, gth, gal, gG, rna, rnr = ao.uode_value_and_grad(
            net.weights, net.biases, net.spec.activation, net.spec.final_activation, net.max_time,
            ts, y0, loss_fn, mech.states, mech.parameters, mech.total_exprs(), net.parameters,
            net.alpha_residual, net.gate_matrix, rtol=1e-7, atol=1e-9)
        assert np.array_equal(net.last["n_accepted"].cpu().numpy(), rna)
        assert np.array_equal(net.last["n_rejected"].cpu().numpy(), rnr)
        np.testing.assert_allclose(net.last["ys"].cpu().numpy(), rys, rtol=1e-9, atol=1e-11)
        assert abs(val - rv) <= 1e-10 * abs(rv)
        _check(g, rW, rb, 1e-7)
        for name, ref in (("parameters", gth), ("alpha_residual", gal), ("gate_matrix", gG)):
            assert np.abs(g[name] - ref).max() <= 1e-7 * np.abs(ref).max(), (name, g[name], ref)
    finally:
        ctx.enable_x64(False)


@pytest.mark.parametrize("activation,fb", [("tanh", False), ("celu", True)])
def test_dopri5_weight_gradient_matches_autograd_of_the_discrete_scheme_f64(activation, fb):
    """trainer.py:128-132 lets the caller pick the solver: the adjoint kernels with Dopri5 steps and
    Shampine's dense output (written as y + dt sum_i w_i(theta) f_i) vs the torch-autograd oracle."""
    import catalax_b200 as ctx
    from catalax_b200.solvers import Dopri5
    from oracle import adjoint_oracle as ao

    ctx.enable_x64(True)
    try:
        net, ts, y0, data = _problem(3, 32, 2, activation, 6, 9, seed=13, use_final_bias=fb)
        val, g = net.value_and_grad(ts, y0, data, want_y0=True, rtol=1e-7, atol=1e-9, solver=Dopri5)
        d = torch.tensor(data)
        B, T, S = data.shape
        loss_fn = lambda ys: (0.5 * (ys - d) ** 2).sum() / (T * B * S)
        rv, rys, rW, rb, rgy0, rna, rnr = ao.neuralode_value_and_grad(
            net.weights, net.biases, net.spec.activation, net.spec.final_activation, net.max_time, ts, y0,
            loss_fn, rtol=1e-7, atol=1e-9, solver="dopri5")
        assert rna.min() >= 3 and rnr.max() >= 1
        assert np.array_equal(net.last["n_accepted"].cpu().numpy(), rna)
        assert np.array_equal(net.last["n_rejected"].cpu().numpy(), rnr)
        np.testing.assert_allclose(net.last["ys"].cpu().numpy(), rys, rtol=1e-9, atol=1e-11)
        assert abs(val - rv) <= 1e-10 * abs(rv)
        _check(g, rW, rb, 1e-8)
        assert np.abs(g["y0"].cpu().numpy() - rgy0).max() <= 1e-8 * np.abs(rgy0).max()
        # the Dopri5 training solve equals the Dopri5 forward solve of the same network
        pred = net.predict_arrays(ts, y0, solver=Dopri5, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(pred, rys, rtol=1e-8, atol=1e-10)
        # ... and differs from Tsit5's step sequence
        v5, _ = net.value_and_grad(ts, y0, data, rtol=1e-7, atol=1e-9)
        assert not np.array_equal(net.last["n_accepted"].cpu().numpy(), rna) or v5 != val
    finally:
        ctx.enable_x64(False)


def test_strategy_curriculum_trains_a_neural_ode_on_the_reference_dataset():
    """examples/NeuralODE.ipynb cells 3-5 in spirit: ``NeuralODE.from_model(model, width_size=4,
    depth=2, activation=celu)`` trained on the reference's Michaelis-Menten trajectories with a
    two-step ``Strategy`` (short horizon first, then the full length, AdaBelief, L2 penalty,
    temporal dropout) through ``model.train(dataset, strategy)``."""
    from pathlib import Path

    import catalax_b200 as ctx
    from catalax_b200.neural import Modes, NeuralODE, Penalties, Strategy

    g = np.load(Path(__file__).parent / "golden" / "croissant_menten.npz")
    ctx.enable_x64(False)
    model = ctx.Model(name="Michaelis-Menten")
    model.add_state(s1="Substrate")
    sub = slice(0, 40)
    ds = ctx.Dataset.from_jax_arrays(["s1"], g["data"][sub, :, None], np.tile(g["times"], (40, 1)),
                                     g["data"][sub, :1])
    net = NeuralODE.from_model(model, width_size=4, depth=2, activation="celu", seed=1)
    assert net.state_order == ["s1"] and net.spec.activation == "celu"
    strategy = Strategy(penalties=Penalties.for_neural_ode(l2_alpha=1e-5), batch_size=20)
    strategy.add_step(lr=1e-2, steps=150, length=0.15)
    strategy.add_step(lr=1e-2, steps=250, length=1.0, train=Modes.MLP)
    before = float(np.mean(np.abs(net.predict_arrays(ds.to_jax_arrays(["s1"])[1], g["data"][sub, :1]) - g["data"][sub, :, None])))
    trained = net.train(ds, strategy, weight_scale=1e-2, seed=3, temporal_dropout_p=0.1, print_every=100)
    # max_time starts at times.max() = 100 and is then TRAINED like the reference's array leaf
    # (trainer.py:85-89, strategy.py:62)
    assert trained is net and 50.0 < net.max_time < 200.0 and abs(net.max_time - 100.0) > 1e-6
    pred = net.predict_arrays(np.tile(g["times"], (40, 1)), g["data"][sub, :1])
    after = float(np.mean(np.abs(pred - g["data"][sub, :, None])))
    print(f"mae before {before:.2f} -> after {after:.2f}")
    # an untrained field leaves the substrate where it started (mae ~ the consumed amount)
    assert after < 0.35 * before, (before, after)
    assert len(net.history) == 400 and np.isfinite(net.history).all()


def test_rateflow_and_universal_ode_train_with_their_penalty_presets():
    """``Penalties.for_rateflow`` / ``for_universal_ode`` inside the curriculum: the stoichiometric
    matrix, the gate matrix and the residual scale get the kernel's data gradient PLUS the host
    autograd gradient of the penalties, and the penalised loss goes down."""
    import catalax_b200 as ctx
    from catalax_b200.neural import Modes, Penalties, RateFlowODE, Strategy, UniversalODE

    ctx.enable_x64(False)
    rng = np.random.default_rng(5)
    M, T = 24, 12
    times = np.tile(np.linspace(0, 5, T), (M, 1))
    a0, b0 = rng.uniform(0.5, 2.0, M), rng.uniform(0.0, 0.5, M)
    k = 0.6
    a = a0[:, None] * np.exp(-k * times)
    b = b0[:, None] + a0[:, None] * (1 - np.exp(-k * times))
    data = np.stack([a, b], axis=-1)

    model = ctx.Model(name="decay")
    model.add_state(a="A", b="B")
    model.add_ode("a", "-k1 * a")
    model.add_ode("b", "k1 * a")
    model.parameters["k1"].value = 0.3
    ds = ctx.Dataset.from_jax_arrays(["a", "b"], data, times, data[:, 0, :])

    rf = RateFlowODE.from_model(model, width_size=8, depth=1, reaction_size=2, seed=2)
    s0 = rf.stoich_matrix.copy()
    pen = Penalties.for_rateflow(alpha=1e-3, conservation_alpha=1e-3)
    strat = Strategy(penalties=pen, batch_size=8).add_step(lr=5e-3, steps=40, length=1.0)
    rf.train(ds, strat, weight_scale=1e-1, seed=1, print_every=1000)
    assert np.isfinite(rf.history).all() and len(rf.history) == 40
    assert np.abs(rf.stoich_matrix - s0).max() > 1e-4          # learn_stoich: the matrix moved
    assert np.mean(rf.history[-5:]) < np.mean(rf.history[:5])

    uo = UniversalODE.from_model(model, width_size=8, depth=1, seed=3)
    g0, a_res0, p0 = uo.gate_matrix.copy(), uo.alpha_residual.copy(), uo.parameters.copy()
    strat = Strategy(penalties=Penalties.for_universal_ode(), batch_size=8)
    strat.add_step(lr=5e-3, steps=30, length=1.0, train=Modes.BOTH)
    uo.train(ds, strat, weight_scale=1e-1, seed=1, print_every=1000)
    assert np.isfinite(uo.history).all()
    assert np.abs(uo.gate_matrix - g0).max() > 0 and np.abs(uo.alpha_residual - a_res0).max() > 0
    assert np.abs(uo.parameters - p0).max() > 0                # Modes.BOTH also fits the mechanistic k1
    assert np.mean(uo.history[-5:]) < np.mean(uo.history[:5])


@pytest.mark.parametrize("kind", ["neuralode", "rateflow"])
def test_max_time_gradient_matches_autograd(kind):
    """The reference's trainer replaces ``func.max_time`` by the array ``times.max()``
    (trainer.py:85-89), which makes it a trainable leaf of the MLP (strategy.py:62).  Its gradient
    comes out of the backward kernel (cotangent of the network's time input) and equals autograd
    through the discrete scheme; the curriculum then moves it like every other leaf."""
    import catalax_b200 as ctx
    from catalax_b200.neural import RateFlowODE
    from oracle import adjoint_oracle as ao

    ctx.enable_x64(True)
    try:
        net, ts, y0, data = _problem(3, 16, 2, "tanh", 6, 9, seed=11)
        stoich = None
        if kind == "rateflow":
            net = RateFlowODE(3, 4, 16, 2, ["a", "b", "c"], activation="celu", max_time=float(ts.max()), key=4)
            net.stoich_matrix = net.stoich_matrix * 0.05      # mild dynamics (see the RateFlow test above)
            stoich = net.stoich_matrix
        val, g = net.value_and_grad(ts, y0, data, rtol=1e-7, atol=1e-9)
        d = torch.tensor(data)
        loss_fn = lambda ys: (0.5 * (ys - d) ** 2).mean()
        ref = ao.neuralode_value_and_grad(net.weights, net.biases, net.spec.activation,
                                          net.spec.final_activation, net.max_time, ts, y0, loss_fn,
                                          rtol=1e-7, atol=1e-9, stoich=stoich, want_max_time=True)
        # (relu kinks: an accept / reject decision within rounding of the threshold may differ
        #  between the two implementations and moves that trajectory by O(rtol))
        rna, rnr = ref[5], ref[6]
        same_path = np.array_equal(net.last["n_accepted"].cpu().numpy(), rna) and \
            np.array_equal(net.last["n_rejected"].cpu().numpy(), rnr)
        tol = 1e-9 if same_path else 2e-6
        assert abs(val - ref[0]) <= 10 * tol * abs(ref[0])
        want = ref[-1]
        assert want != 0.0 and abs(g["max_time"] - want) <= 100 * tol * abs(want), (g["max_time"], want)
        # ... and a finite difference of the loss in max_time
        h = 1e-5 * net.max_time
        keep = net.max_time
        vals = []
        for mt in (keep + h, keep - h):
            net.max_time = mt
            net._compiled.clear()
            vals.append(net.value_and_grad(ts, y0, data, rtol=1e-7, atol=1e-9)[0])
        net.max_time = keep
        net._compiled.clear()
        fd = (vals[0] - vals[1]) / (2 * h)
        # (the finite difference also sees the step sequence move with max_time; the gradient is
        #  the derivative on the frozen sequence: agreement to the solver's own accuracy)
        assert abs(fd - want) <= (1e-4 if kind == "neuralode" else 1e-2) * abs(want), (fd, want)
    finally:
        ctx.enable_x64(False)
