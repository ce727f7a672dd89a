"""BASELINE.json's FULL sizes on the GPU, checked through size-independent properties of the
domain (the oracle only sees subsamples it finishes in seconds):

* configs[1] (2^20 Michaelis-Menten rows x 1000 saved points, f32): the enzyme is conserved
  exactly, P + S is conserved to rounding, S decreases monotonically, the first saved point is
  y0, every row succeeds; a random subsample agrees with the closed form (Lambert W) and with
  the C oracle at the f32 bound;
* configs[3] (16384 chains x 64 series): the log-likelihood block is additive over series
  (two halves sum to the whole) and chains are independent (a chain evaluated alone gives the
  same bits);
* configs[2] / configs[4] at full batch: every row succeeds, outputs finite, a subsample agrees
  with the oracle / the x64 FFMA form.
"""
import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_config2_full_size_conservation_monotonicity_and_closed_form():
    from scipy.special import lambertw

    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle import c_oracle

    B, T = 1 << 20, 1000
    dev = torch.device("cuda:0")
    y0, th, c = cases.mm_inputs(B, seed=0, dtype=np.float32)
    model = cases.michaelis_menten()
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), "float32")
    tt = lambda a: torch.as_tensor(a, device=dev)
    ts = np.linspace(0, 100, T).astype(np.float32)
    o = m.solve(tt(y0), tt(th), tt(c), tt(ts), in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6,
                dt0=0.1, max_steps=4096)
    torch.cuda.synchronize()
    ys = o.ys                                                    # [B,T,3] = (E, P, S), 12.6 GB
    assert int((o.status != 0).sum().item()) == 0
    assert bool(torch.isfinite(ys).all().item())
    y0d = tt(y0)
    assert bool((ys[:, 0, :] == y0d).all().item())               # ts[0] = t0: y0 itself
    assert bool((ys[:, :, 0] == y0d[:, None, 0]).all().item())   # dE/dt = 0: exact
    total = (y0d[:, 1] + y0d[:, 2])[:, None]
    drift = ((ys[:, :, 1] + ys[:, :, 2]) - total).abs() / total
    assert float(drift.max().item()) <= 5e-6                     # P + S conserved (f32 rounding)
    dS = ys[:, 1:, 2] - ys[:, :-1, 2]
    assert float(dS.max().item()) <= 2e-5 * float(y0d[:, 2].max().item())   # S never increases
    del drift, dS
    # subsample: closed form S(t) = K W((S0/K) exp((S0 - V t)/K)), V = k_cat E0
    idx = np.random.default_rng(0).choice(B, 2048, replace=False)
    sub = ys[torch.as_tensor(idx, device=dev)].cpu().numpy().astype(np.float64)
    K, V = th[idx, 0].astype(np.float64), (th[idx, 1] * y0[idx, 0]).astype(np.float64)
    S0 = y0[idx, 2].astype(np.float64)
    arg = (S0 / K)[:, None] * np.exp((S0[:, None] - V[:, None] * ts[None, :].astype(np.float64)) / K[:, None])
    S_exact = K[:, None] * np.real(lambertw(arg))
    err = np.abs(sub[:, :, 2] - S_exact) / (1e-6 + 1e-6 * np.abs(S_exact))
    assert np.percentile(err, 99.9) <= 200 and np.median(err) <= 20, (np.median(err), err.max())
    # ... and the C oracle (the reference algorithm in f32) on the same rows, same bound
    ref = c_oracle.COracle(*model[:4], model[4], dtype=np.float32).solve(
        y0[idx], th[idx], c[idx], ts, rtol=1e-6, atol=1e-6, dt0=0.1, max_steps=4096)
    e2 = np.abs(sub - ref.ys.astype(np.float64)) / (1e-6 + 1e-6 * np.abs(ref.ys))
    assert np.percentile(e2, 99.9) <= 200
    steps_gpu = (o.n_accepted + o.n_rejected)[torch.as_tensor(idx, device=dev)].cpu().numpy()
    steps_ref = ref.n_accepted + ref.n_rejected
    assert abs(steps_gpu.sum() / steps_ref.sum() - 1.0) < 0.02   # same algorithm, same work


def test_config4_full_size_additivity_over_series_and_chain_independence():
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    CH, M, T = 16384, 64, 20
    rng = np.random.default_rng(2)
    dev = torch.device("cuda:0")
    model = cases.michaelis_menten()
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model), sens_theta=True), "float32")
    y0s = np.stack([rng.uniform(5, 20, M), np.zeros(M), rng.uniform(50, 300, M)], 1).astype(np.float32)
    ts = np.tile(np.linspace(0, 100, T), (M, 1)).astype(np.float32)
    data = (y0s[:, None, :] * rng.uniform(0.2, 1.0, (M, T, 3))).astype(np.float32)
    data[rng.uniform(size=data.shape) < 0.05] = np.nan
    theta = (np.array([40.0, 0.3]) * np.exp(rng.normal(0, 0.2, (CH, 2)))).astype(np.float32)
    sigma = np.exp(rng.normal(np.log(0.5), 0.3, (CH, 3))).astype(np.float32)
    consts = np.zeros((M, 0), np.float32)
    tt = lambda a: torch.as_tensor(a, device=dev)

    def run(ms, chains=slice(None)):
        out, _, st = m.loglik_grad(tt(y0s[ms]), tt(theta[chains]), tt(consts[ms]), tt(ts[ms]),
                                   tt(data[ms]), tt(sigma[chains]), likelihood="softlaplace")
        torch.cuda.synchronize()
        assert int((st != 0).sum().item()) == 0
        return out.cpu().numpy().astype(np.float64)

    whole = run(slice(0, M))
    assert whole.shape == (CH, 6) and np.isfinite(whole).all()
    halves = run(slice(0, M // 2)) + run(slice(M // 2, M))
    scale = np.abs(whole) + 1e-3 * np.abs(whole).max(axis=1, keepdims=True)
    assert (np.abs(whole - halves) <= 2e-5 * scale).all(), float((np.abs(whole - halves) / scale).max())
    some = np.array([0, 1, 777, 9000, CH - 1])
    np.testing.assert_array_equal(run(slice(0, M), some), whole[some])   # chains do not interact


def test_config3_and_config5_full_batches_succeed_and_match_on_a_subsample():
    from catalax_b200 import engine
    from catalax_b200.neural import NeuralODE
    from catalax_b200.tools import codegen as cg
    from oracle import c_oracle

    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    # configs[2]: 32 states / 64 reactions, 262144 rows
    B3 = 262144
    net3 = cases.mass_action_network(32)
    rng = np.random.default_rng(1)
    y3 = rng.uniform(0.1, 2.0, (B3, 32)).astype(np.float32)
    k3 = np.exp(rng.uniform(np.log(0.01), np.log(1.0), (B3, 64))).astype(np.float32)
    ts3 = np.linspace(0, 10, 32).astype(np.float32)
    m3 = engine.CompiledModel(cg.generate_field(cases.field_spec(net3)), "float32")
    o = m3.solve(tt(y3), tt(k3), tt(np.zeros((B3, 0), np.float32)), tt(ts3), in_axes=(0, 0, 0, None),
                 rtol=1e-6, atol=1e-6)
    torch.cuda.synchronize()
    assert int((o.status != 0).sum().item()) == 0 and bool(torch.isfinite(o.ys).all().item())
    assert bool((o.ys >= -1e-4).all().item())                     # concentrations stay non-negative
    idx = np.random.default_rng(3).choice(B3, 256, replace=False)
    ref = c_oracle.COracle(*net3[:4], net3[4], dtype=np.float64).solve(
        y3[idx].astype(np.float64), k3[idx].astype(np.float64), np.zeros((256, 0)), ts3.astype(np.float64),
        rtol=1e-9, atol=1e-9)
    sub = o.ys[tt(idx)].cpu().numpy().astype(np.float64)
    err = np.abs(sub - ref.ys) / (1e-6 + 1e-6 * np.abs(ref.ys))
    assert np.percentile(err, 99.9) <= 200, float(err.max())
    # configs[4]: NeuralODE 5-64-64-64-4 on the tensor cores, 524288 rows
    B5 = 1 << 19
    net = NeuralODE(4, 64, 3, ["a", "b", "c", "d"], activation="tanh", max_time=10.0, key=3)
    y5 = np.random.default_rng(3).uniform(0, 2, (B5, 4)).astype(np.float32)
    ts5 = np.linspace(0, 10, 16).astype(np.float32)
    ys5 = net.predict_arrays(tt(ts5), tt(y5))
    torch.cuda.synchronize()
    assert int((net.last.status != 0).sum().item()) == 0 and bool(torch.isfinite(ys5).all().item())
    import catalax_b200 as ctx
    ctx.enable_x64(True)
    try:
        want = net.predict_arrays(tt(ts5.astype(np.float64)), tt(y5[:4096].astype(np.float64))).cpu().numpy()
    finally:
        ctx.enable_x64(False)
    got = ys5[:4096].cpu().numpy().astype(np.float64)
    e5 = np.abs(got - want) / (1e-6 + 1e-3 * np.abs(want))        # rtol 1e-3, atol 1e-6 (neuralbase.py:634)
    assert (e5 <= 20).mean() >= 0.99 and np.median(e5) <= 1.0
