"""``reconstruct_log_likelihood`` / ``integrate_ensemble`` on the GPU (one fused two-level launch each)
against the arrays the reference's own loo.py / predictive.py functions produce
(tests/golden/make_posterior_callers_golden.py)."""
import types

import numpy as np
import pytest

from tests.test_posterior_callers_golden import G, POSTERIOR

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _results(lik):
    import catalax_b200 as ctx

    model = ctx.Model(name="MM")
    model.add_state(S="Substrate", E="Enzyme", P="Product")
    model.add_ode("S", "-k_cat * E * S / (K_m + S)")
    model.add_ode("P", "k_cat * E * S / (K_m + S)")
    model.add_ode("E", "0")
    from catalax_b200.mcmc import MCMCConfig

    bm = types.SimpleNamespace(priors=[("K_m", None), ("k_cat", None)], likelihood=lik, mode="integrated",
                               config=MCMCConfig(0, 0).to_simulation_config())
    return types.SimpleNamespace(bayesian_model=bm, model=model,
                                 mcmc=types.SimpleNamespace(get_samples=lambda group_by_chain=True: POSTERIOR))


@pytest.mark.parametrize("lik", ["normal", "softlaplace"])
def test_reconstruct_log_likelihood_equals_the_reference_functions(lik):
    import catalax_b200 as ctx
    from catalax_b200.mcmc import loo

    ctx.enable_x64(True)
    try:
        res = _results(lik)
        arrays = (G["data"], G["times"], G["y0s"], np.zeros((G["y0s"].shape[0], 0)))
        for leave_out in ("point", "curve"):
            for max_draws in (None, 3):
                ll, post, meta = loo.reconstruct_log_likelihood(res, None, leave_out=leave_out,
                                                                max_draws=max_draws, arrays=arrays)
                key = f"{lik}/{leave_out}/{max_draws}"
                want = G[f"{key}/loglik"]
                np.testing.assert_array_equal(meta["kept_idx"], G[f"{key}/kept_idx"])
                np.testing.assert_array_equal(post["K_m"], G[f"{key}/post_K_m"])
                assert ll.shape == want.shape
                # x64: the solve agrees with the oracle's to ~1e-12
                assert np.abs(ll - want).max() <= 1e-8 * np.abs(want).max()
        ll, _, _ = loo.reconstruct_log_likelihood(res, None, sigma_source="yerrs",
                                                  yerrs=np.array([0.5, 1.0, 2.0]), max_draws=3, arrays=arrays)
        want = G[f"{lik}/yerrs/loglik"]
        assert np.abs(ll - want).max() <= 1e-8 * np.abs(want).max()
    finally:
        ctx.enable_x64(False)


def test_integrate_ensemble_equals_the_reference_function():
    import catalax_b200 as ctx
    from catalax_b200.mcmc import predictive as pr

    ctx.enable_x64(True)
    try:
        res = _results("normal")
        arrays = (G["data"], G["times"], G["y0s"], np.zeros((G["y0s"].shape[0], 0)))
        sub = pr.subsample(POSTERIOR, 4)
        times, yhat = pr.integrate_ensemble(res, None, sub, n_steps=17, arrays=arrays)
        np.testing.assert_allclose(times, G["ens/times"], rtol=1e-14)
        want = G["ens/yhat"]
        assert yhat.shape == want.shape == (12, 5, 17, 3)
        assert np.abs(yhat - want).max() <= 1e-9 * np.abs(want).max()
    finally:
        ctx.enable_x64(False)


@pytest.mark.parametrize("lik", ["normal", "softlaplace"])
def test_surrogate_mode_reconstruction_equals_the_reference_functions(lik):
    """loo.py:203-303: rates of every draw at the measured states (ctx_rates), Euler-integrated and
    scored -- against the reference's `_surrogate_loglik` + `euler_integrate_rates` executed over
    the oracle's rate function."""
    import catalax_b200 as ctx
    from catalax_b200.mcmc import loo

    ctx.enable_x64(True)
    try:
        res = _results(lik)
        res.bayesian_model.mode = "surrogate"
        M = G["y0s"].shape[0]
        arrays = (G["surr/full_data"], G["times"], None, np.zeros((M, 0)))
        for integration in ("euler", "euler_onestep"):
            ll, _, meta = loo.reconstruct_log_likelihood(res, None, integration=integration, max_draws=4,
                                                         arrays=arrays)
            want = G[f"surr/{lik}/{integration}/loglik"]
            np.testing.assert_array_equal(meta["kept_idx"], G[f"surr/{lik}/{integration}/kept_idx"])
            assert meta["mode"] == "SURROGATE" and ll.shape == want.shape
            assert np.abs(ll - want).max() <= 1e-9 * np.abs(want).max()
        ll, _, _ = loo.reconstruct_log_likelihood(res, None, sigma_source="yerrs", yerrs=1.3, leave_out="curve",
                                                  max_draws=4, arrays=arrays)
        want = G[f"surr/{lik}/yerrs_curve/loglik"]
        assert np.abs(ll - want).max() <= 1e-9 * np.abs(want).max()
        with pytest.raises(ValueError, match="yerrs"):
            loo.reconstruct_log_likelihood(res, None, sigma_source="yerrs", arrays=arrays)
    finally:
        ctx.enable_x64(False)
