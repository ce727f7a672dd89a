"""K2 (fused log-likelihood + gradient) vs the oracle and vs finite differences.

BASELINE configs[3] at a size the oracle finishes in seconds: Michaelis-Menten, chains x series
x 20 points x 3 observables, rtol=atol=1e-5, dt0=0.1, max_steps=4096 (what MCMC really uses,
mcmc.py:73-79), Normal and SoftLaplace likelihoods, NaN-masked points."""
import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


THETA_STAR = (40.0, 0.3)   # K_m, k_cat: non-stiff (decay rate k_cat*E0/K_m <= 0.15 per time unit)
THETA_README = (0.05, 0.1)  # README values: rows deplete and turn stiff -> see the last test


def _config4(CH, M, T=20, seed=2, dtype=np.float64, theta_star=THETA_STAR):
    from oracle import diffrax_oracle as do, symbolic as sy

    rng = np.random.default_rng(seed)
    model = cases.michaelis_menten()
    S0 = rng.uniform(50, 300, M); E0 = rng.uniform(5, 20, M)
    y0s = np.stack([E0, np.zeros(M), S0], axis=1)
    ts = np.tile(np.linspace(0, 100, T), (M, 1))
    theta_star = np.array(theta_star)
    rhs, *_ = sy.make_rhs(*model[:4], model[4])
    clean = do.solve(rhs, y0s, theta_star, np.zeros((M, 0)), ts, rtol=1e-8, atol=1e-8).ys
    data = clean + rng.normal(size=clean.shape) * (0.05 * np.abs(clean) + 1e-3)
    data[rng.uniform(size=data.shape) < 0.05] = np.nan          # missing observations
    theta = theta_star * np.exp(rng.normal(0, 0.2, (CH, 2)))
    sigma = np.exp(rng.normal(np.log(0.5), 0.3, (CH, 3)))
    sigma[0, 1] *= -1.0                                          # scale = sqrt(sigma^2) = |sigma|
    cast = lambda a: a.astype(dtype)
    return model, cast(y0s), cast(theta), np.zeros((M, 0), dtype), cast(ts), cast(data), cast(sigma)


def _gpu(model, y0s, theta, consts, ts, data, sigma, likelihood, dtype, sens_y0=False, **kw):
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    field = cg.generate_field(cases.field_spec(model), sens_theta=True, sens_y0=sens_y0)
    m = engine.CompiledModel(field, np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=dtype), device=dev)
    out, partial, status = m.loglik_grad(tt(y0s), tt(theta), tt(consts), tt(ts), tt(data), tt(sigma),
                                         likelihood=likelihood, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy(), partial.cpu().numpy(), status.cpu().numpy()


@pytest.mark.parametrize("likelihood", ["normal", "softlaplace"])
def test_loglik_grad_matches_oracle_f64(likelihood):
    from oracle import loglik_oracle as lo

    CH, M = 12, 16
    model, y0s, theta, consts, ts, data, sigma = _config4(CH, M)
    out, partial, status = _gpu(model, y0s, theta, consts, ts, data, sigma, likelihood, np.float64)
    ref, _, rstat = lo.loglik_grad(model, y0s, theta, consts, ts, data, sigma,
                                   likelihood=likelihood, rtol=1e-5, atol=1e-5)
    assert (status == 0).all() and (rstat == 0).all()
    # x64 bound: 1e-6 relative (plus 1e-9 of the row's largest |gradient| for cancelling sums)
    scale = np.abs(ref) + 1e-3 * np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(out - ref) <= 1e-6 * scale).all(), float((np.abs(out - ref) / scale).max())
    assert out.shape == (CH, 1 + 2 + 3)


def test_loglik_grad_y0_directions_f64():
    from oracle import loglik_oracle as lo

    CH, M = 4, 8
    model, y0s, theta, consts, ts, data, sigma = _config4(CH, M, seed=5)
    out, partial, status = _gpu(model, y0s, theta, consts, ts, data, sigma, "normal", np.float64,
                                sens_y0=True)
    ref, gy0, _ = lo.loglik_grad(model, y0s, theta, consts, ts, data, sigma, likelihood="normal",
                                 sens_y0=True, rtol=1e-5, atol=1e-5)
    assert partial.shape == (CH * M, 1 + 2 + 3 + 3)
    got = partial[:, -3:].reshape(CH, M, 3)
    scale = np.abs(gy0) + 1e-3 * np.abs(gy0).max()
    assert (np.abs(got - gy0) <= 1e-6 * scale).all()
    np.testing.assert_allclose(partial[:, 0].reshape(CH, M).sum(1), out[:, 0], rtol=1e-12)


def test_gradient_equals_finite_differences_of_loglik():
    """d/dsigma exactly; d/dtheta up to the O(tol) step-size-freezing term (stop_gradient)."""
    CH, M = 3, 8
    model, y0s, theta, consts, ts, data, sigma = _config4(CH, M, seed=7)
    kw = dict(rtol=1e-10, atol=1e-10, max_steps=100000)
    out, _, _ = _gpu(model, y0s, theta, consts, ts, data, sigma, "softlaplace", np.float64, **kw)
    eps = 1e-6
    for k in range(2):
        tp, tm = theta.copy(), theta.copy()
        tp[:, k] *= 1 + eps; tm[:, k] *= 1 - eps
        lp, _, _ = _gpu(model, y0s, tp, consts, ts, data, sigma, "softlaplace", np.float64, **kw)
        lm, _, _ = _gpu(model, y0s, tm, consts, ts, data, sigma, "softlaplace", np.float64, **kw)
        fd = (lp[:, 0] - lm[:, 0]) / (2 * eps * theta[:, k])
        np.testing.assert_allclose(out[:, 1 + k], fd, rtol=2e-5)
    for k in range(3):
        sp_, sm_ = sigma.copy(), sigma.copy()
        sp_[:, k] += eps; sm_[:, k] -= eps
        lp, _, _ = _gpu(model, y0s, theta, consts, ts, data, sp_, "softlaplace", np.float64, **kw)
        lm, _, _ = _gpu(model, y0s, theta, consts, ts, data, sm_, "softlaplace", np.float64, **kw)
        np.testing.assert_allclose(out[:, 3 + k], (lp[:, 0] - lm[:, 0]) / (2 * eps), rtol=1e-6)


def test_loglik_f32_stated_bound():
    """f32 bound: |diff| <= 2e-3 * (|ref| + max|row|*1e-2) vs the x64 oracle at the same tolerances."""
    from oracle import loglik_oracle as lo

    CH, M = 8, 16
    model, y0s, theta, consts, ts, data, sigma = _config4(CH, M, seed=9, dtype=np.float32)
    out, _, status = _gpu(model, y0s, theta, consts, ts, data, sigma, "normal", np.float32)
    ref, _, _ = lo.loglik_grad(model, y0s.astype(np.float64), theta.astype(np.float64),
                               consts.astype(np.float64), ts.astype(np.float64),
                               data.astype(np.float64), sigma.astype(np.float64),
                               likelihood="normal", rtol=1e-5, atol=1e-5)
    assert (status == 0).all()
    scale = np.abs(ref) + 1e-2 * np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(out - ref) <= 2e-3 * scale).all(), float((np.abs(out - ref) / scale).max())


def test_failed_solve_gives_minus_inf():
    """throw=False path (mcmc.py:216-219): a failed series poisons the chain's log-density."""
    CH, M = 2, 4
    model, y0s, theta, consts, ts, data, sigma = _config4(CH, M, seed=11)
    out, _, status = _gpu(model, y0s, theta, consts, ts, data, sigma, "normal", np.float64,
                          rtol=1e-9, atol=1e-9, max_steps=12)
    assert (status == 1).any()
    assert np.isneginf(out[:, 0]).all()


def test_readme_parameters_discrete_tangents_are_ill_conditioned_but_consistent():
    """With the README values (K_m=0.05) most series deplete before t=100 and the explicit solver
    becomes stability-limited on a component far below atol.  The error controller does not see
    the tangents (they are outside the norm, diffrax's stop_gradient semantics), so the DISCRETE
    sensitivities grow to ~1e7 and the gradient is a sum of huge cancelling terms -- a property
    of the reference's algorithm, not of this implementation.  The log-likelihood still agrees to
    1e-6; the gradient agrees relative to the L1 norm of its summands."""
    from oracle import diffrax_oracle as do, loglik_oracle as lo, symbolic as sy

    CH, M = 4, 8
    model, y0s, theta, consts, ts, data, sigma = _config4(CH, M, seed=2, theta_star=THETA_README)
    out, _, status = _gpu(model, y0s, theta, consts, ts, data, sigma, "normal", np.float64)
    ref, _, _ = lo.loglik_grad(model, y0s, theta, consts, ts, data, sigma, likelihood="normal",
                               rtol=1e-5, atol=1e-5)
    assert (status == 0).all()
    np.testing.assert_allclose(out[:, 0], ref[:, 0], rtol=1e-6)
    np.testing.assert_allclose(out[:, 3:], ref[:, 3:], rtol=1e-6)
    rhs, N, D, ia = sy.make_rhs(*model[:4], model[4], sens_theta=True)
    for ch in range(CH):
        r = do.solve(rhs, y0s, theta[ch], consts, ts, rtol=1e-5, atol=1e-5, n_err=3,
                     init_aug=ia(M, np.float64))
        tang = r.ys[..., 3:].reshape(M, 20, 2, 3)
        mask = np.isfinite(data)
        z = (np.where(mask, data, 0.0) - r.ys[..., :3]) / np.abs(sigma[ch]) ** 2
        l1 = np.abs(np.where(mask, z, 0.0)[:, :, None, :] * tang).sum(axis=(0, 1, 3))
        # rounding-level differences are amplified ~1e10x by the unstable tangent recursion
        assert (np.abs(out[ch, 1:3] - ref[ch, 1:3]) <= 5e-3 * l1).all()


def test_bayesian_model_log_density_through_the_facade():
    """BayesianModel.log_density = priors + error-model prior + fused CUDA likelihood block
    (mcmc.py:405-491), built from a Model + Dataset exactly like HMC.run does (:170-297)."""
    import catalax_b200 as ctx
    from catalax_b200.mcmc import BayesianModel, error, priors
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
        CH, M = 6, 10
        mm, y0s, theta, consts, ts, data, sigma = _config4(CH, M, seed=21)
        sigma = np.abs(sigma)
        ds = ctx.Dataset.from_jax_arrays(model.get_state_order(), data, ts, y0s)
        bm = BayesianModel(model, ds, yerrs=0.5, likelihood="normal",
                           error_model=error.LogNormal(median=0.5))
        dev = torch.device("cuda:0")
        th_t, sg_t = torch.as_tensor(theta, device=dev), torch.as_tensor(sigma, device=dev)
        logp, g_th, g_sg = bm.log_density(th_t, sg_t)
        torch.cuda.synchronize()
        ref, _, _ = lo.loglik_grad(mm, y0s, theta, consts, ts, data, sigma, likelihood="normal",
                                   rtol=1e-5, atol=1e-5)
        th_c, sg_c = torch.as_tensor(theta), torch.as_tensor(sigma)
        want = torch.as_tensor(ref[:, 0]) + priors.Uniform(1.0, 200.0).log_prob(th_c[:, 0]) \
            + priors.LogUniform(0.01, 10.0).log_prob(th_c[:, 1]) \
            + error.LogNormal(median=0.5).log_prob(sg_c, torch.full_like(sg_c, 0.5))
        np.testing.assert_allclose(logp.cpu().numpy(), want.numpy(), rtol=1e-7)
        want_gth = ref[:, 1:3].copy()
        want_gth[:, 1] += priors.LogUniform(0.01, 10.0).grad(th_c[:, 1]).numpy()
        np.testing.assert_allclose(g_th.cpu().numpy(), want_gth, rtol=2e-6)
        want_gsg = ref[:, 3:] + error.LogNormal(median=0.5).grad(sg_c, torch.full_like(sg_c, 0.5)).numpy()
        np.testing.assert_allclose(g_sg.cpu().numpy(), want_gsg, rtol=2e-6)
    finally:
        ctx.enable_x64(False)


def test_posterior_ensemble_two_level_batch():
    """compute_ensemble's vmap(per_draw)(thetas) (predictive.py:150-154): [draws, M, T, S] from one
    launch with ctx_solve_cfg.inner = M; compared with per-draw oracle solves, and its point-wise
    log-likelihood summed over the data equals the fused K2 kernel's."""
    import catalax_b200 as ctx
    from catalax_b200.mcmc.predictive import EnsembleIntegrator, ensemble_times
    from oracle import diffrax_oracle as do, symbolic as sy

    ctx.enable_x64(True)
    try:
        model = ctx.Model(name="MM")
        model.add_state(S="Substrate", E="Enzyme", P="Product")
        model.add_ode("S", "-k_cat * E * S / (K_m + S)")
        model.add_ode("P", "k_cat * E * S / (K_m + S)")
        model.add_ode("E", "0")
        D, M = 7, 9
        mm, y0s, theta, consts, ts, data, sigma = _config4(D, M, seed=31)
        grid = ensemble_times(ts, 33)
        ens = EnsembleIntegrator(model)
        ys = ens.integrate(y0s, theta, consts, grid).cpu().numpy()
        assert ys.shape == (D, M, 33, 3)
        rhs, *_ = sy.make_rhs(*mm[:4], mm[4])
        for d in range(D):
            ref = do.solve(rhs, y0s, theta[d], consts, grid, rtol=1e-5, atol=1e-5, dtype=np.float64)
            np.testing.assert_allclose(ys[d], ref.ys, rtol=1e-6, atol=1e-9)
        lp = ens.pointwise_log_likelihood(y0s, theta, sigma, consts, ts, data, "normal").cpu().numpy()
        out, _, _ = _gpu(mm, y0s, theta, consts, ts, data, sigma, "normal", np.float64)
        np.testing.assert_allclose(np.nansum(lp, axis=(1, 2, 3)), out[:, 0], rtol=1e-9)
    finally:
        ctx.enable_x64(False)


@pytest.mark.parametrize("kernel", [1, 2])
def test_both_loglik_kernels_match_oracle(kernel):
    """ctx_loglik_v1 (per-lane terms) and ctx_loglik_v2 (warp-cooperative terms) are the same
    computation: both meet the x64 bound against the oracle, incl. the y0 directions."""
    from oracle import loglik_oracle as lo

    CH, M = 9, 21   # ragged: 189 trajectories, not a multiple of the warp size
    model, y0s, theta, consts, ts, data, sigma = _config4(CH, M, seed=11)
    out, partial, status = _gpu(model, y0s, theta, consts, ts, data, sigma, "softlaplace",
                                np.float64, sens_y0=True, kernel=kernel)
    ref, gy0, rstat = lo.loglik_grad(model, y0s, theta, consts, ts, data, sigma,
                                     likelihood="softlaplace", rtol=1e-5, atol=1e-5, sens_y0=True)
    assert (status == 0).all() and (rstat == 0).all()
    scale = np.abs(ref) + 1e-3 * np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(out - ref) <= 1e-6 * scale).all(), float((np.abs(out - ref) / scale).max())
    g = partial.reshape(CH, M, -1)[:, :, 1 + 2 + 3:]
    gs = np.abs(gy0) + 1e-3 * np.abs(gy0).max()
    assert (np.abs(g - gy0) <= 1e-6 * gs).all()


def test_cooperative_loglik_is_reproducible_and_handles_long_segments():
    """Run-to-run bit reproducibility of the cooperative kernel (time-ordered accumulation), on a
    grid where one step covers many requested times (segments longer than a warp pass) and on
    one where t0 == t1 for every point."""
    CH, M, T = 64, 8, 150
    model, y0s, theta, consts, ts, data, sigma = _config4(CH, M, T=T, seed=5)
    a = _gpu(model, y0s, theta, consts, ts, data, sigma, "normal", np.float32, kernel=2)
    b = _gpu(model, y0s, theta, consts, ts, data, sigma, "normal", np.float32, kernel=2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # same grid in x64: both kernels agree far below the solver tolerance (they differ only in the
    # algebraic form of the interpolant and in the summation order)
    v2 = _gpu(model, y0s, theta, consts, ts, data, sigma, "normal", np.float64, kernel=2)
    v1 = _gpu(model, y0s, theta, consts, ts, data, sigma, "normal", np.float64, kernel=1)
    scale = np.abs(v1[0]) + 1e-3 * np.abs(v1[0]).max(axis=1, keepdims=True)
    assert (np.abs(v2[0] - v1[0]) <= 1e-8 * scale).all(), float((np.abs(v2[0] - v1[0]) / scale).max())
    assert (v2[2] == 0).all() and (v1[2] == 0).all()
    # degenerate grid: all requested times are t0 = 0
    ts0 = np.zeros_like(ts)
    z2 = _gpu(model, y0s, theta, consts, ts0, data, sigma, "normal", np.float64, kernel=2)
    z1 = _gpu(model, y0s, theta, consts, ts0, data, sigma, "normal", np.float64, kernel=1)
    assert np.allclose(z2[0], z1[0], rtol=1e-12, atol=1e-12)
    assert (z2[2] == 0).all()


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_order_hint_leaves_the_log_density_bits_unchanged(kernel, dtype):
    """The scheduling hint of ctx_loglik_grad (trajectories handed out most-expensive-first, from
    the step counts of the previous evaluation -- what a sampler has at every leapfrog step): any
    permutation gives the same [CH, 1+P+S] block, partial rows and status, bit for bit."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    CH, M = 150, 13
    model, y0s, theta, consts, ts, data, sigma = _config4(CH, M, dtype=dtype)
    field = cg.generate_field(cases.field_spec(model), sens_theta=True)
    m = engine.CompiledModel(field, np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=dtype), device=dev)
    args = [tt(x) for x in (y0s, theta, consts, ts, np.nan_to_num(data), sigma)]
    mask = torch.as_tensor(np.isfinite(data), device=dev)
    base = m.loglik_grad(*args, mask=mask, kernel=kernel)
    order = engine.cost_order(*m.last_loglik_counts)
    assert sorted(order.tolist()) == list(range(CH * M))
    rnd = torch.as_tensor(np.random.default_rng(1).permutation(CH * M).astype(np.int32), device=dev)
    for o in (order, rnd):
        got = m.loglik_grad(*args, mask=mask, kernel=kernel, order=o)
        torch.cuda.synchronize()
        for a_, b_ in zip(got, base):
            assert torch.equal(a_, b_)
    with pytest.raises(ValueError, match="order must be"):
        m.loglik_grad(*args, mask=mask, order=order[:5].contiguous())
