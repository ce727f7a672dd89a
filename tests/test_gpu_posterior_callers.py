"""SURVEY §8 f1 -- the posterior re-integration callers on the GPU: the fused point-wise
log-likelihood (ctx_pointwise_loglik; catalax/mcmc/loo.py:131-200), the ensemble integration
(catalax/mcmc/predictive.py:109-155, 212-249) and the ``HMC.run`` / ``run_mcmc`` entry points
(catalax/mcmc/mcmc.py:170-293, 494-538) that produce the posterior they consume."""
import numpy as np
import pytest

import catalax_b200 as ctx
from catalax_b200 import workloads as wl

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _problem(D=37, M=11, T=20, seed=3):
    from oracle.c_oracle import COracle

    mm = wl.michaelis_menten()
    clean = lambda y0s, th, ts: COracle(*mm[:4], mm[4], dtype=np.float64).solve(
        y0s, np.tile(th, (len(y0s), 1)), np.zeros((len(y0s), 0)), ts, rtol=1e-9, atol=1e-9,
        max_steps=100000).ys
    y0s, theta, consts, ts, data, sigma = wl.nuts_operands(D, M, T, seed, clean)
    rng = np.random.default_rng(seed)
    ts = ts + rng.uniform(0, 0.3, (M, 1)) * np.arange(T)[None, :] / T      # per-series grids
    return mm, y0s, theta, consts, ts, data, sigma


@pytest.mark.parametrize("likelihood", ["normal", "softlaplace"])
@pytest.mark.parametrize("T", [20, 67])
def test_fused_pointwise_loglik_matches_the_oracle_f64(likelihood, T):
    """ll[d,m,t,o] from ONE fused launch == likelihood(loc = oracle trajectory of (draw d, series
    m), scale = |sigma[d,o]|).log_prob(data[m,t,o]) (loo.py:180-193), 1e-6 relative in x64."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle import loglik_oracle as lo
    from oracle.c_oracle import COracle

    mm, y0s, theta, consts, ts, data, sigma = _problem(T=T)
    D, M = theta.shape[0], y0s.shape[0]
    sigma[0, 1] *= -1.0          # only |sigma| enters
    m = engine.CompiledModel(cg.generate_field(wl.field_spec(mm)), "float64")
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)
    out = m.pointwise_loglik(tt(y0s), tt(theta), tt(consts), tt(ts), tt(data), tt(sigma),
                             likelihood=likelihood, rtol=1e-6, atol=1e-6, inner=M)
    torch.cuda.synchronize()
    assert engine.last_solve_kernel().startswith("ctx_solve_v1 (log-likelihood save)")
    ll = out.ys.cpu().numpy().reshape(D, M, T, 3)
    orc = COracle(*mm[:4], mm[4], dtype=np.float64)
    ref = orc.solve(np.tile(y0s, (D, 1)), np.repeat(theta, M, axis=0), np.tile(consts, (D, 1)),
                    np.tile(ts, (D, 1)), rtol=1e-6, atol=1e-6)
    assert (out.status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    loc = ref.ys.reshape(D, M, T, 3)
    want, _, _ = lo.logpdf_terms(data[None], loc, np.abs(sigma)[:, None, None, :], likelihood)
    assert np.isfinite(ll).all()
    assert np.abs(ll - want).max() <= 1e-6 * np.abs(want).max(), float(np.abs(ll - want).max())
    # ... and the same numbers as evaluating the likelihood on the materialised trajectories
    ys = m.solve(tt(y0s), tt(theta), tt(consts), tt(ts), in_axes=(0, 0, 0, 0), rtol=1e-6, atol=1e-6,
                 inner=M).ys.cpu().numpy().reshape(D, M, T, 3)
    two_pass, _, _ = lo.logpdf_terms(data[None], ys, np.abs(sigma)[:, None, None, :], likelihood)
    assert np.abs(ll - two_pass).max() <= 1e-9 * np.abs(want).max()


def test_fused_pointwise_loglik_f32_and_failed_rows():
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle import loglik_oracle as lo

    mm, y0s, theta, consts, ts, data, sigma = _problem(D=64, M=8)
    m = engine.CompiledModel(cg.generate_field(wl.field_spec(mm)), "float32")
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)
    out = m.pointwise_loglik(tt(y0s), tt(theta), tt(consts), tt(ts), tt(data), tt(sigma),
                             likelihood="normal", inner=8)
    ys = m.solve(tt(y0s), tt(theta), tt(consts), tt(ts), in_axes=(0, 0, 0, 0), inner=8).ys
    want, _, _ = lo.logpdf_terms(data[None].astype(np.float32), ys.cpu().numpy().reshape(64, 8, 20, 3),
                                 np.abs(sigma).astype(np.float32)[:, None, None, :], "normal")
    got = out.ys.cpu().numpy().reshape(64, 8, 20, 3)
    assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()
    # a draw whose solve fails (max_steps) scores -inf on everything it never reached (inf states)
    out2 = m.pointwise_loglik(tt(y0s), tt(theta), tt(consts), tt(ts), tt(data), tt(sigma),
                              likelihood="normal", inner=8, max_steps=3)
    assert int((out2.status != 0).sum()) > 0
    tail = out2.ys.reshape(64, 8, 20, 3)[:, :, -1, 1:]
    assert bool(torch.isinf(tail).all()) and bool((tail < 0).all())
    with pytest.raises(ValueError, match="finite"):
        bad = data.copy(); bad[0, 0, 0] = np.nan
        m.pointwise_loglik(tt(y0s), tt(theta), tt(consts), tt(ts), tt(bad), tt(sigma), inner=8)


def _menten_posterior_problem():
    from catalax_b200.mcmc import priors

    ctx.enable_x64(False)
    model = ctx.Model(name="Simple menten model")
    model.add_state("s1")
    model.add_ode("s1", "- (v_max * s1) / ( K_m + s1)")
    model.parameters["v_max"].value = 7.0
    model.parameters["K_m"].value = 100.0
    ds = ctx.Dataset.from_model(model)
    for s0 in (60.0, 120.0, 200.0, 300.0):
        ds.add_initial(s1=s0)
    sim = model.simulate(ds, ctx.SimulationConfig(t0=0, t1=100, nsteps=20))
    noisy = sim.augment(n_augmentations=1, sigma=0.03, seed=1, append=False, multiplicative=True)
    model.parameters["v_max"].prior = priors.Uniform(low=1e-6, high=200.0)
    model.parameters["K_m"].prior = priors.Uniform(low=1e-6, high=1e3)
    return model, noisy


def test_hmc_run_then_loo_and_ensemble_callers():
    """End to end through the reference's entry-point names: ``HMC(...).run(model, dataset,
    yerrs)`` samples the posterior (every leapfrog step = one fused log-density launch for all
    chains), then ``mechanistic_loglik`` (LOO input) and ``compute_ensemble`` (predictive bands)
    re-integrate the model per draw."""
    from catalax_b200.mcmc import HMC, MCMCConfig, loo, predictive, run_mcmc

    model, data = _menten_posterior_problem()
    hmc = HMC(num_warmup=150, num_samples=100, num_chains=64, max_tree_depth=3, verbose=0, seed=7)
    res = hmc.run(model=model, dataset=data, yerrs=2.0)
    post = res.mcmc.get_samples(group_by_chain=True)
    assert post["K_m"].shape == (64, 100) and post["sigma_y"].shape == (64, 100, 1)
    flat = res.get_samples()
    assert 0.4 < res.stats["mean_accept_prob"] <= 1.0
    # the posterior concentrates on the parameters that generated the data
    assert abs(np.median(flat["v_max"]) - 7.0) < 0.7 and abs(np.median(flat["K_m"]) - 100.0) < 30.0
    # run_mcmc = HMC.from_config(config).run(...) (mcmc.py:494-538)
    cfg = MCMCConfig(num_warmup=20, num_samples=5, num_chains=8, max_tree_depth=2, verbose=0)
    assert run_mcmc(model, data, 2.0, cfg).get_samples()["v_max"].shape == (40,)

    sub = predictive.subsample(post, 10)
    ll, mask, info = loo.mechanistic_loglik(res, data, sub)
    assert ll.shape == (64, 10, 4, 20, 1) and mask.shape == (4, 20, 1) and np.isfinite(ll).all()
    # the fused kernel's numbers equal the likelihood of the per-draw trajectories
    integ = predictive.EnsembleIntegrator(model, loo.build_eval_config(res))
    thetas, _ = predictive.stack_thetas(sub, ["K_m", "v_max"])
    from catalax_b200.mcmc.mcmc import extract_dataset_components
    d_, t_, y0_, c_ = extract_dataset_components(data, model)
    ys = integ.integrate(y0_, thetas, c_, t_).cpu().numpy()
    sig = predictive.sigma_y_draws(sub, 64, 10, 1)[:, None, None, :]
    z = (d_[None] - ys) / sig
    want = np.log(2 / np.pi) - np.log(sig) - np.logaddexp(z, -z)            # SoftLaplace (default)
    assert np.abs(ll.reshape(want.shape) - want).max() < 2e-3
    elpd = loo.pointwise_elpd_is(ll, mask)
    assert elpd.shape == mask.shape and np.isfinite(elpd).all()

    times, yhat, var = predictive.compute_ensemble(res, data, n_steps=30, max_draws=5)
    assert times.shape == (4, 30) and yhat.shape == (64 * 5, 4, 30, 1) == var.shape
    assert np.isfinite(yhat).all() and (np.diff(yhat[..., 0], axis=2) <= 1e-3).all()   # substrate decays


def test_bayesian_model_accepts_the_reference_call_order():
    """``BayesianModel.__call__(y0s, constants, times, mask, data)`` (mcmc.py:405) binds the data
    arguments and returns the log-density over them."""
    from catalax_b200.mcmc import BayesianModel
    from catalax_b200.mcmc.mcmc import extract_dataset_components

    model, data = _menten_posterior_problem()
    bm = BayesianModel(model, data, yerrs=2.0)
    dev = torch.device("cuda:0")
    theta = torch.tensor([[100.0, 7.0], [90.0, 6.5]], device=dev)
    sigma = torch.tensor([[2.0], [1.5]], device=dev).expand(2, 1).contiguous()
    lp0, g0, s0 = bm(theta, sigma)
    d_, t_, y0_, c_ = extract_dataset_components(data, model)
    fn = bm(y0_, c_, t_, np.isfinite(d_), np.nan_to_num(d_))
    lp1, g1, s1 = fn(theta, sigma)
    assert torch.equal(lp0, lp1) and torch.equal(g0, g1)
    mask = np.isfinite(d_); mask[0, 5:] = False         # drop observations: the density changes
    lp2, _, _ = bm(y0_, c_, t_, mask, np.nan_to_num(d_))(theta, sigma)
    assert not torch.equal(lp0, lp2)
