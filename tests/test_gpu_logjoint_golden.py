"""``BayesianModel.log_density`` (CUDA: K2 / K5 + host priors) against the log-joint the reference's
own ``BayesianModel`` class computes (tests/golden/make_logjoint_golden.py), integrated and
surrogate mode, every error model, both likelihoods, NaN-masked and ragged data."""
import numpy as np
import pytest

from tests import cases  # noqa: F401  (path set-up)
from tests.test_logjoint_golden import G, PRIOR_SETS, error_model_of, sigma_of, yerrs_of

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _model(pset_name):
    import catalax_b200 as ctx
    from catalax_b200.mcmc import priors

    model = ctx.Model(name="MM")
    model.add_state(S="Substrate", E="Enzyme", P="Product")
    model.add_ode("S", "-k_cat * E * S / (K_m + S)")
    model.add_ode("P", "k_cat * E * S / (K_m + S)")
    model.add_ode("E", "0")
    for name, prior in PRIOR_SETS[pset_name](priors).items():
        model.parameters[name].prior = prior
    assert model.get_state_order() == list(G["states"]) and model.get_parameter_order() == list(G["params"])
    return model


@pytest.mark.parametrize("case", list("ABCDEFG"))
def test_log_density_equals_the_reference_log_joint(case):
    import catalax_b200 as ctx
    from catalax_b200.mcmc import BayesianModel

    ctx.enable_x64(True)
    try:
        pset, lik = str(G[f"{case}/meta"][0]), str(G[f"{case}/meta"][1])
        model = _model(pset)
        M, T = G["times"].shape
        kw = dict(yerrs=yerrs_of(case), likelihood=lik, error_model=error_model_of(case), shard="none")
        if case in "FG":
            extra2 = None
            if str(G[f"{case}/meta"][3]) == "True":
                extra2 = G["surr_sigma_surr"] ** 2 + G["surr_rate_sigma"][None, :] ** 2
            y = G["surr_y"].reshape(M, T, 3)
            bm = BayesianModel(model, arrays=(y, G["times"], y[:, 0, :], np.zeros((M, 0))),
                               surrogate_arrays=(G["surr_rates"], G["surr_jac"], extra2), **kw)
        else:
            bm = BayesianModel(model, arrays=(G["data"], G["times"], G["y0s"], np.zeros((M, 0))), **kw)
        dev = torch.device("cuda:0")
        theta = torch.as_tensor(G["theta"], device=dev)
        sigma = torch.as_tensor(sigma_of(case), device=dev)
        logp, g_theta, g_sigma = bm.log_density(theta, sigma)
        want = sum(G[k] for k in G.files if k.startswith(f"{case}/") and k.split("/")[1] in
                   ("K_m", "k_cat", "sigma_y", "y"))
        # x64: the solve agrees with the oracle's to ~1e-12; 1e-9 of the log-joint
        np.testing.assert_allclose(logp.cpu().numpy(), want, rtol=1e-9)
        assert g_theta.shape == (5, 2) and g_sigma.shape == sigma.shape
        assert torch.isfinite(g_theta).all() and torch.isfinite(g_sigma).all()
    finally:
        ctx.enable_x64(False)
