"""The NUTS log-joint against values produced by the reference's own source lines
(tests/golden/make_logjoint_golden.py: catalax/mcmc/priors.py, error.py and the complete
``BayesianModel`` class of mcmc.py executed unmodified over numpy stand-ins for jax / numpyro).

CPU half: the host-side pieces (priors, error models, scale hint) and the parity oracle's
likelihood blocks (integrated and surrogate mode).  The GPU half is in
tests/test_gpu_logjoint_golden.py."""
from pathlib import Path

import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
G = np.load(Path(__file__).parent / "golden" / "logjoint_reference_lines.npz")

PRIOR_SETS = {
    "uniform_loguniform": lambda pr: {"K_m": pr.Uniform(1.0, 200.0), "k_cat": pr.LogUniform(0.01, 10.0)},
    "normal_truncnormal": lambda pr: {"K_m": pr.Normal(45.0, 12.0), "k_cat": pr.TruncatedNormal(0.25, 0.2)},
}


def error_model_of(case):
    from catalax_b200.mcmc import error

    return {"A": None, "B": error.LogNormal(sigma=0.7), "C": error.Gamma(concentration=2.0, shared=True),
            "D": error.Fixed(np.array([0.3, 0.9, 1.1])), "E": error.HalfNormal(scale=1.7),
            "F": None, "G": error.LogNormal(sigma=0.5)}[case]


def sigma_of(case):
    if case == "C":
        return G["sigma1"]
    if case == "D":
        return np.broadcast_to(np.array([0.3, 0.9, 1.1]), G["sigma3"].shape).copy()
    return G["sigma3"]


def yerrs_of(case):
    return G[f"{case}/yerrs"] if f"{case}/yerrs" in G.files else np.broadcast_to(0.8, G["data"].shape)


@pytest.mark.parametrize("case", list("ABCDEFG"))
def test_priors_match_the_reference_sites(case):
    from catalax_b200.mcmc import priors as pr

    pset = PRIOR_SETS[str(G[f"{case}/meta"][0])](pr)
    theta = torch.as_tensor(G["theta"])
    for j, name in enumerate(["K_m", "k_cat"]):           # the reference's (sorted) parameter order
        got = pset[name].log_prob(theta[:, j]).numpy()
        np.testing.assert_allclose(got, G[f"{case}/{name}"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("case", list("ABCEFG"))
def test_error_models_match_the_reference_sigma_y_site(case):
    from catalax_b200.mcmc import error
    from catalax_b200.mcmc.mcmc import per_species_scale_hint

    em = error_model_of(case) or error.DEFAULT_ERROR_MODEL
    sigma = torch.as_tensor(sigma_of(case))
    hint = torch.as_tensor(per_species_scale_hint(yerrs_of(case), 3))
    if getattr(em, "shared", False):
        hint = hint.mean().reshape(1)
    got = em.log_prob(sigma, hint.expand_as(sigma)).numpy()
    np.testing.assert_allclose(got, G[f"{case}/sigma_y"], rtol=1e-12, atol=1e-12)


def test_fixed_error_model_has_no_sampled_site():
    assert "D/sigma_y" not in G.files and not error_model_of("D").sampled


@pytest.mark.parametrize("case", list("ABCDE"))
def test_oracle_likelihood_block_matches_the_reference_y_site(case):
    """loglik_oracle (what every CUDA log-likelihood test is compared with) == the reference's
    masked ``sample("y", likelihood(loc, sqrt(sigma_total)), obs=data)`` on the same states."""
    from oracle import loglik_oracle as lo

    model = cases.michaelis_menten()
    lik = str(G[f"{case}/meta"][1])
    sigma = np.broadcast_to(sigma_of(case), G["sigma3"].shape)
    M = G["y0s"].shape[0]
    out, _, status = lo.loglik_grad(model, G["y0s"], G["theta"], np.zeros((M, 0)), G["times"], G["data"],
                                    sigma, likelihood=lik, rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096)
    assert (status == 0).all()
    np.testing.assert_allclose(out[:, 0], G[f"{case}/y"], rtol=1e-12)


@pytest.mark.parametrize("case", ["F", "G"])
def test_oracle_surrogate_block_matches_the_reference_y_site(case):
    from oracle import loglik_oracle as lo

    model = cases.michaelis_menten()
    lik = str(G[f"{case}/meta"][1])
    M, T = G["times"].shape
    y = G["surr_y"]
    extra2 = None
    if str(G[f"{case}/meta"][3]) == "True":                # mcmc.py:474-477
        extra2 = G["surr_sigma_surr"] ** 2 + G["surr_rate_sigma"][None, :] ** 2
    out = lo.surrogate_loglik_grad(model, G["surr_t"], y, G["theta"], np.zeros((M, 0)),
                                   y.reshape(M, T, 3)[:, 0, :], G["surr_rates"], G["surr_jac"] ** 2,
                                   G["sigma3"], n_time=T, extra2=extra2, likelihood=lik)
    np.testing.assert_allclose(out[:, 0], G[f"{case}/y"], rtol=1e-12)
