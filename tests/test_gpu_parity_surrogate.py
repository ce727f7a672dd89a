"""K5 (surrogate / rate-matching log-likelihood + gradient) vs the oracle.

Reference path: catalax/mcmc/mcmc.py:819-847 (rate function at the measured states),
:469-480 (sigma_y pushed forward through the surrogate's rate-Jacobian), :486-491 (masked
likelihood).  The reference asserts no value of this density anywhere (parity unpinned against
numpyro); pinned here against the numpy restatement and finite differences."""
import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _problem(CH=5, M=7, T=9, seed=3, dtype=np.float64, with_extra=True):
    rng = np.random.default_rng(seed)
    model = cases.michaelis_menten()
    S = 3
    # measured states along plausible trajectories, grouped per measurement
    E = rng.uniform(5, 20, (M, 1)).repeat(T, 1)
    S_ = rng.uniform(50, 300, (M, 1)) * np.exp(-np.linspace(0, 2, T))[None, :]
    P_ = rng.uniform(0, 5, (M, 1)) + (S_[:, :1] - S_)
    y = np.stack([E, P_, S_], axis=-1).reshape(M * T, S)
    t = np.tile(np.linspace(0, 100, T), M)
    theta = np.array([40.0, 0.3]) * np.exp(rng.normal(0, 0.2, (CH, 2)))
    sigma = np.exp(rng.normal(np.log(0.5), 0.3, (CH, S)))
    sigma[0, 2] *= -1.0                      # only sigma^2 enters
    rate_true = 0.3 * y[:, 0] * y[:, 2] / (40.0 + y[:, 2])
    data = np.stack([np.zeros(M * T), rate_true, -rate_true], axis=1)
    data = data + rng.normal(size=data.shape) * 0.05
    data[rng.uniform(size=data.shape) < 0.07] = np.nan
    jac = rng.normal(size=(M * T, S, S)) * 0.1
    jac2 = jac ** 2
    extra2 = rng.uniform(0.0, 0.01, (M * T, S)) if with_extra else None
    inits = y.reshape(M, T, S)[:, 0, :].copy()
    consts = np.zeros((M, 0))
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dtype)
    return model, dict(t=c(t), y=c(y), theta=c(theta), consts=c(consts), inits=c(inits),
                       data=c(data), jac2=c(jac2), sigma=c(sigma), extra2=c(extra2), n_time=T)


def _gpu(model, pr, likelihood, dtype):
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    field = cg.generate_field(cases.field_spec(model), sens_theta=True)
    m = engine.CompiledModel(field, np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=dtype), device=dev)
    out = m.rate_loglik_grad(tt(pr["t"]), tt(pr["y"]), tt(pr["theta"]), tt(pr["consts"]),
                             tt(pr["inits"]), tt(np.nan_to_num(pr["data"])), tt(pr["jac2"]),
                             tt(pr["sigma"]), n_time=pr["n_time"],
                             mask=torch.as_tensor(np.isfinite(pr["data"]), device=dev),
                             extra2=tt(pr["extra2"]), likelihood=likelihood)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("likelihood", ["normal", "softlaplace"])
@pytest.mark.parametrize("with_extra", [True, False])
def test_rate_loglik_matches_oracle_f64(likelihood, with_extra):
    from oracle import loglik_oracle as lo

    model, pr = _problem(with_extra=with_extra)
    out = _gpu(model, pr, likelihood, np.float64)
    ref = lo.surrogate_loglik_grad(model, **pr, likelihood=likelihood)
    assert out.shape == ref.shape == (5, 1 + 2 + 3)
    scale = np.abs(ref) + 1e-6 * np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(out - ref) <= 1e-9 * scale).all(), float((np.abs(out - ref) / scale).max())


def test_rate_loglik_gradient_equals_finite_differences():
    from oracle import loglik_oracle as lo

    model, pr = _problem(CH=2, seed=8)
    out = _gpu(model, pr, "softlaplace", np.float64)
    for col, (key, j) in enumerate([("theta", 0), ("theta", 1), ("sigma", 0), ("sigma", 1), ("sigma", 2)]):
        h = 1e-6 * np.abs(pr[key][:, j])
        hi = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pr.items()}
        lo_ = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pr.items()}
        hi[key][:, j] += h
        lo_[key][:, j] -= h
        fd = (lo.surrogate_loglik_grad(model, **hi)[:, 0] - lo.surrogate_loglik_grad(model, **lo_)[:, 0]) / (2 * h)
        np.testing.assert_allclose(out[:, 1 + col], fd, rtol=2e-6, atol=1e-7)


def test_rate_loglik_f32_stated_bound():
    """f32 (the reference's default dtype): 2e-4 of the row scale against the x64 oracle."""
    from oracle import loglik_oracle as lo

    model, pr = _problem(CH=16, M=12, T=20, dtype=np.float32)
    out = _gpu(model, pr, "softlaplace", np.float32)
    ref = lo.surrogate_loglik_grad(model, **pr, likelihood="softlaplace")
    scale = np.abs(ref) + 1e-2 * np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(out - ref) <= 2e-4 * scale).all(), float((np.abs(out - ref) / scale).max())


def test_init_symbols_resolve_to_the_first_state_of_each_measurement():
    """`{state}_init` in a rate law = the measurement's first measured state (mcmc.py:833-842)."""
    from oracle import loglik_oracle as lo

    import sympy as sp

    model = (["p", "s"], ["k"], [], {"p": sp.sympify("k * s_init"), "s": sp.sympify("-k * s")}, [])
    rng = np.random.default_rng(1)
    M, T, S, CH = 4, 5, 2, 3
    y = rng.uniform(0.5, 2.0, (M * T, S))
    pr = dict(t=np.tile(np.linspace(0, 3, T), M), y=y, theta=rng.uniform(0.5, 1.5, (CH, 1)),
              consts=np.zeros((M, 0)), inits=y.reshape(M, T, S)[:, 0, :].copy(),
              data=rng.normal(size=(M * T, S)), jac2=rng.uniform(0, 1, (M * T, S, S)),
              sigma=rng.uniform(0.3, 1.0, (CH, S)), extra2=None, n_time=T)
    out = _gpu(model, pr, "normal", np.float64)
    ref = lo.surrogate_loglik_grad(model, **pr, likelihood="normal")
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-10)


def test_bayesian_model_surrogate_mode_through_the_facade():
    """BayesianModel(surrogate=NeuralODE): rates and Jacobian come from the surrogate's public API
    (K4 kernels), the density from K5; checked against the oracle fed with the same arrays."""
    import catalax_b200 as ctx
    from catalax_b200.mcmc import BayesianModel, priors
    from catalax_b200.mcmc.mcmc import surrogate_rate_jacobian
    from catalax_b200.neural import NeuralODE
    from oracle import loglik_oracle as lo

    ctx.enable_x64(True)
    try:
        model = ctx.Model(name="MM")
        model.add_state(S="Substrate", E="Enzyme", P="Product")
        model.add_ode("S", "-k_cat * E * S / (K_m + S)")
        model.add_ode("P", "k_cat * E * S / (K_m + S)")
        model.add_ode("E", "0")
        model.parameters.K_m.prior = priors.Uniform(1.0, 200.0)
        model.parameters.k_cat.prior = priors.LogUniform(0.01, 10.0)
        rng = np.random.default_rng(0)
        M, T, CH = 5, 8, 4
        y0s = np.stack([rng.uniform(5, 20, M), np.zeros(M), rng.uniform(50, 300, M)], 1)
        times = np.tile(np.linspace(0, 60, T), (M, 1))
        data = np.stack([np.broadcast_to(y0s[:, :1], (M, T)),
                         y0s[:, 2:3] * (1 - np.exp(-times / 40.0)),
                         y0s[:, 2:3] * np.exp(-times / 40.0)], axis=-1)
        net = NeuralODE(3, 16, 2, model.get_state_order(), activation="tanh", max_time=60.0, key=1)
        pts = data.reshape(M * T, 3)
        rates = np.asarray(net.rates(times.ravel(), pts), dtype=np.float64)
        jac = surrogate_rate_jacobian(net, pts, times.ravel())
        assert jac.shape == (M * T, 3, 3)
        # Jacobian sanity: a directional finite difference of the public rates() API
        v = rng.normal(size=3)
        h = 1e-5
        fd = (np.asarray(net.rates(times.ravel(), pts + h * v)) - np.asarray(net.rates(times.ravel(), pts - h * v))) / (2 * h)
        np.testing.assert_allclose(jac @ v, fd, rtol=1e-5, atol=1e-7)
        bm = BayesianModel(model, arrays=(data, times, y0s, np.zeros((M, 0))), likelihood="normal",
                           shard="none", surrogate_arrays=(rates, jac, None))
        dev = torch.device("cuda:0")
        theta = torch.tensor(np.array([40.0, 0.3]) * np.exp(rng.normal(0, 0.1, (CH, 2))), device=dev)
        sigma = torch.tensor(rng.uniform(0.3, 2.0, (CH, 3)), device=dev)
        block = bm.log_likelihood(theta, sigma).cpu().numpy()
        mm = cases.michaelis_menten()
        ref = lo.surrogate_loglik_grad(mm, times.ravel(), pts, theta.cpu().numpy(), np.zeros((M, 0)),
                                       data[:, 0, :], rates, jac ** 2, sigma.cpu().numpy(), n_time=T,
                                       likelihood="normal")
        np.testing.assert_allclose(block, ref, rtol=1e-9, atol=1e-9)
        logp, g_theta, g_sigma = bm.log_density(theta, sigma)
        assert logp.shape == (CH,) and g_theta.shape == (CH, 2) and g_sigma.shape == (CH, 3)
        assert torch.isfinite(logp).all()
    finally:
        ctx.enable_x64(False)
