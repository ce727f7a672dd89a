"""Posterior re-integration callers against arrays produced by the reference's own
functions (tests/golden/make_posterior_callers_golden.py executes loo.py:92-200, 306-437 and
predictive.py:109-175 unmodified over numpy stand-ins).  CPU half: the host bookkeeping around the
fused kernel; the GPU half is in tests/test_gpu_posterior_callers_golden.py."""
from pathlib import Path

import numpy as np
import pytest

G = np.load(Path(__file__).parent / "golden" / "posterior_callers_reference_lines.npz")
POSTERIOR = {"K_m": G["post_K_m"], "k_cat": G["post_k_cat"], "sigma_y": G["post_sigma_y"]}


@pytest.mark.parametrize("lik", ["normal", "softlaplace"])
@pytest.mark.parametrize("leave_out", ["point", "curve"])
def test_leave_out_bookkeeping(lik, leave_out):
    from catalax_b200.mcmc import loo

    loglik, meta = loo.leave_out_groups(G[f"{lik}/full"], G["mask"], leave_out)
    key = f"{lik}/{leave_out}/None"
    np.testing.assert_array_equal(meta["kept_idx"], G[f"{key}/kept_idx"])
    assert meta["n_dropped"] == int(G[f"{key}/n_dropped"])
    np.testing.assert_allclose(loglik, G[f"{key}/loglik"], rtol=0, atol=0)
    with pytest.raises(ValueError, match="leave_out"):
        loo.leave_out_groups(G[f"{lik}/full"], G["mask"], "series")


def test_draw_helpers():
    from catalax_b200.mcmc import predictive as pr

    sub = pr.subsample(POSTERIOR, 3)
    np.testing.assert_array_equal(sub["K_m"], G["normal/point/3/post_K_m"])
    assert pr.subsample(POSTERIOR, None) is POSTERIOR and pr.subsample(POSTERIOR, 100) is POSTERIOR
    sub4 = pr.subsample(POSTERIOR, 4)
    np.testing.assert_array_equal(sub4["K_m"], G["ens/sub_K_m"])
    thetas, (n_chain, n_draw) = pr.stack_thetas(sub4, ["K_m", "k_cat"])
    assert (n_chain, n_draw) == (3, 4) and thetas.shape == (12, 2)
    np.testing.assert_array_equal(thetas[:, 0], sub4["K_m"].reshape(-1))
    sig = pr.sigma_y_draws(sub4, 3, 4, 3)
    np.testing.assert_array_equal(sig, sub4["sigma_y"].reshape(12, 3))
    np.testing.assert_array_equal(pr.sigma_y_draws({}, 3, 4, 3), np.ones((12, 3)))
    np.testing.assert_allclose(pr.ensemble_times(G["times"], 17), G["ens/times"], rtol=1e-15)
    np.testing.assert_allclose(pr.aleatoric_variance(sub4, G["ens/yhat"].shape), G["ens/var"], rtol=1e-15)


@pytest.mark.parametrize("mode", ["euler", "euler_onestep"])
def test_euler_integrate_rates(mode):
    """predictive.py:54-106, single draw and with a leading draw axis."""
    from catalax_b200.mcmc.predictive import euler_integrate_rates

    kw = dict(integration=mode, data_safe=G["euler/data_safe"])
    yhat, var_y = euler_integrate_rates(G["euler/rates"], G["euler/y0_obs"], G["euler/dt"], G["euler/rate_var"], **kw)
    np.testing.assert_allclose(yhat, G[f"euler/{mode}/yhat"], rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(var_y, G[f"euler/{mode}/var_y"], rtol=1e-14, atol=1e-14)
    stacked = np.stack([G["euler/rates"], 2.0 * G["euler/rates"]])
    yhat2, _ = euler_integrate_rates(stacked, G["euler/y0_obs"], G["euler/dt"], G["euler/rate_var"], **kw)
    np.testing.assert_allclose(yhat2[0], yhat, rtol=1e-14, atol=1e-14)
    assert yhat2.shape == (2,) + yhat.shape


@pytest.mark.parametrize("hdi", [None, "lower", "upper", "lower_50", "upper_50"])
@pytest.mark.parametrize("noise", [True, False])
def test_band_values(hdi, noise):
    """predictive.py:252-293: posterior-median trajectory and percentile bands, with the aleatoric
    noise of PRNGKey(0) pooled in."""
    from catalax_b200.mcmc.predictive import band_values

    got = band_values(G["ens/yhat"], G["ens/var"], hdi=hdi, include_noise=noise)
    np.testing.assert_allclose(got, G[f"band/{hdi}/{noise}"], rtol=1e-6, atol=1e-6)
    with pytest.raises(KeyError):
        band_values(G["ens/yhat"], G["ens/var"], hdi="median")
