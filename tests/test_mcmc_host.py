"""Host-side logic of the log-density entry point (no GPU): priors, error models, sharding and
the world_size-2 all-reduce of the likelihood block (gloo)."""
import math
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]

from catalax_b200 import parallel
from catalax_b200.mcmc import error, priors


@pytest.mark.parametrize("prior", [priors.Normal(1.0, 0.5), priors.TruncatedNormal(1.0, 0.5, 0.1, 3.0),
                                   priors.Uniform(0.1, 3.0), priors.LogUniform(0.1, 3.0)])
def test_prior_gradients_match_autograd_and_normalise(prior):
    x = torch.linspace(0.2, 2.5, 7, dtype=torch.float64, requires_grad=True)
    lp = prior.log_prob(x)
    if isinstance(prior, priors.Uniform):
        assert (prior.grad(x.detach()) == 0).all()
    else:
        (g,) = torch.autograd.grad(lp.sum(), x)
        np.testing.assert_allclose(prior.grad(x.detach()).numpy(), g.numpy(), rtol=1e-12, atol=1e-12)
    if not isinstance(prior, priors.Normal):
        grid = torch.linspace(0.1, 3.0, 200001, dtype=torch.float64)
        mass = torch.trapezoid(torch.exp(prior.log_prob(grid)), grid).item()
        assert abs(mass - 1.0) < 1e-5
        assert prior.log_prob(torch.tensor([5.0], dtype=torch.float64)).item() == -math.inf


@pytest.mark.parametrize("em", [error.HalfNormal(), error.HalfNormal(scale=0.3), error.LogNormal(),
                                error.LogNormal(median=0.2, shared=True), error.Gamma(), error.Fixed(0.1)])
def test_error_model_gradients_match_autograd(em):
    hint = torch.tensor([0.5, 0.25, 1.0], dtype=torch.float64)
    s = torch.tensor([[0.4, 0.3, 0.9], [1.2, 0.1, 0.5]], dtype=torch.float64, requires_grad=True)
    lp = em.log_prob(s, hint.expand_as(s))
    assert lp.shape == (2,)
    if em.sampled:
        (g,) = torch.autograd.grad(lp.sum(), s)
        np.testing.assert_allclose(em.grad(s.detach(), hint.expand_as(s)).numpy(), g.numpy(), rtol=1e-12)


def test_default_error_model_is_halfnormal_of_mean_yerrs():
    # error.py:181 + mcmc.py:316-329: HalfNormal(scale_hint) per observable
    s = torch.tensor([[0.5, 0.5]], dtype=torch.float64)
    hint = torch.tensor([[2.0, 2.0]], dtype=torch.float64)
    want = 2 * (math.log(2) - math.log(2.0) - 0.5 * math.log(2 * math.pi) - 0.5 * 0.25 ** 2)
    assert abs(error.DEFAULT_ERROR_MODEL.log_prob(s, hint).item() - want) < 1e-14


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


_WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["CTX_ROOT"])
import numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["CTX_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
import catalax_b200 as ctx
from catalax_b200.mcmc import BayesianModel, priors
from oracle import loglik_oracle as lo
from tests import cases

ctx.enable_x64(True)
model = ctx.Model(name="MM")
model.add_state(S="Substrate", E="Enzyme", P="Product")
model.add_ode("S", "-k_cat * E * S / (K_m + S)"); model.add_ode("P", "k_cat * E * S / (K_m + S)")
model.add_ode("E", "0")
model.parameters.K_m.prior = priors.Uniform(1.0, 200.0)
model.parameters.k_cat.prior = priors.LogUniform(0.01, 10.0)
rng = np.random.default_rng(0)
M, T, CH = 7, 6, 3
y0s = np.stack([rng.uniform(5, 20, M), np.zeros(M), rng.uniform(50, 300, M)], 1)
times = np.tile(np.linspace(0, 50, T), (M, 1))
data = rng.uniform(1, 100, (M, T, 3)); data[0, 1, 2] = np.nan
arrays = (data, times, y0s, np.zeros((M, 0)))
theta = torch.tensor(np.array([40.0, 0.3]) * np.exp(rng.normal(0, 0.1, (CH, 2))))
sigma = torch.tensor(rng.uniform(0.3, 2.0, (CH, 3)))
mm = cases.michaelis_menten()

def oracle_block(bm):
    # stands in for the CUDA kernel on this CPU-only box: oracle over THIS rank's series
    def f(th, sg):
        d, t, y, c = bm._arrays
        if d.shape[0] == 0:
            return torch.zeros((th.shape[0], 6), dtype=torch.float64)
        out, _, _ = lo.loglik_grad(mm, y, th.numpy(), c, t, d, sg.numpy(), likelihood="normal",
                                   rtol=1e-5, atol=1e-5)
        return torch.tensor(out)
    return f

full = BayesianModel(model, arrays=arrays, likelihood="normal", shard="none")
full._local_loglik = oracle_block(full)
sharded = BayesianModel(model, arrays=arrays, likelihood="normal", shard="series")
sharded._local_loglik = oracle_block(sharded)
assert sharded._arrays[0].shape[0] in (3, 4)
a = full.log_density(theta, sigma)
b = sharded.log_density(theta, sigma)
for x, y in zip(a, b):
    np.testing.assert_allclose(x.numpy(), y.numpy(), rtol=1e-12, atol=1e-12)
chains = BayesianModel(model, arrays=arrays, likelihood="normal", shard="chains")
assert chains._arrays[0].shape[0] == M
dist.barrier(); dist.destroy_process_group()
print("ok", os.environ["RANK"])
"""


def test_series_sharded_loglik_allreduce_world2_gloo(tmp_path):
    import socket
    import subprocess

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", CTX_PORT=str(port), CTX_ROOT=str(ROOT))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o[-3000:]
        assert f"ok {rank}" in o


def test_surrogate_rate_jacobian_by_central_differences_and_series_slicing():
    """Host side of the surrogate mode (mcmc.py:865-894, 819-847): the Jacobian helper works on
    any object with the public ``rates`` API; surrogate arrays follow the series sharding."""
    import catalax_b200 as ctx
    from catalax_b200.mcmc import BayesianModel
    from catalax_b200.mcmc.mcmc import surrogate_rate_jacobian

    class Fake:
        def rates(self, t, y, constants=None):
            y = np.asarray(y, dtype=np.float64)
            return np.stack([y[:, 0] * y[:, 1], np.sin(y[:, 1]) + np.asarray(t)], axis=1)

    rng = np.random.default_rng(0)
    y = rng.uniform(0.5, 2.0, (11, 2)); t = rng.uniform(0, 1, 11)
    J = surrogate_rate_jacobian(Fake(), y, t)
    want = np.zeros((11, 2, 2))
    want[:, 0, 0] = y[:, 1]; want[:, 0, 1] = y[:, 0]; want[:, 1, 1] = np.cos(y[:, 1])
    np.testing.assert_allclose(J, want, rtol=1e-8, atol=1e-9)

    model = ctx.Model(name="decay")
    model.add_state(s="s", p="p")
    model.add_ode("s", "-k * s"); model.add_ode("p", "k * s_init")
    model.parameters.k.prior = priors.Uniform(0.1, 2.0)
    M, T = 4, 3
    data = rng.uniform(1, 2, (M, T, 2)); times = np.tile(np.linspace(0, 1, T), (M, 1))
    rates = rng.normal(size=(M * T, 2)); jac = rng.normal(size=(M * T, 2, 2))
    bm = BayesianModel(model, arrays=(data, times, data[:, 0, :], np.zeros((M, 0))), shard="none",
                       surrogate_arrays=(rates, jac, None))
    assert bm.mode == "surrogate" and bm._surr[0].shape == (M * T, 2) and bm._surr[1].shape == (M * T, 2, 2)
    np.testing.assert_allclose(bm._surr[1], jac ** 2, rtol=1e-6)
    assert bm.jac_rownorm.shape == (2,)
    with pytest.raises(NotImplementedError):
        BayesianModel(model, arrays=(data, times, data[:, 0, :], np.zeros((M, 0))), shard="none",
                      surrogate_arrays=(rates, jac, None), estimate_initials=True)


# ----------------------------------------------------------------------------------------------
# The reference's tests/unit/test_error_model.py, against the facade: how ``yerrs`` becomes the
# per-observable default scale and how each error model turns it into its prior.  The reference
# inspects the numpyro distribution handed to ``sample``; here the prior's log-density is compared
# with scipy's for the scale the reference asserts.
HINT = np.array([0.1, 1.0, 5.0])
N_OBS = 3


def test_scale_hint_from_yerrs():
    from catalax_b200.mcmc.mcmc import per_species_scale_hint

    np.testing.assert_allclose(per_species_scale_hint(0.5, N_OBS), np.full(N_OBS, 0.5))       # :39-44
    np.testing.assert_allclose(per_species_scale_hint(HINT, N_OBS), HINT)                     # :46-49
    full = np.stack([np.full((4, N_OBS), 0.2), np.full((4, N_OBS), 0.4)]) * np.array([1.0, 2.0, 3.0])
    np.testing.assert_allclose(per_species_scale_hint(full, N_OBS),                           # :51-64
                               full.reshape(-1, N_OBS).mean(axis=0))
    odd = np.array([0.1, 0.2, 0.3, 0.4, 0.5])                                                 # :66-71
    np.testing.assert_allclose(per_species_scale_hint(odd, N_OBS), np.full(N_OBS, odd.mean()))


def _log_prior(em, sigma, hint=HINT):
    s = torch.as_tensor(np.atleast_2d(sigma), dtype=torch.float64)
    h = torch.as_tensor(np.ascontiguousarray(hint), dtype=torch.float64)
    if getattr(em, "shared", False):
        h = h.mean().reshape(1)
    return em.log_prob(s, h.expand_as(s)).item()


def test_error_model_scale_is_per_species_or_shared_or_explicit():
    from scipy import stats

    sig = np.array([0.07, 0.8, 3.0])
    # :102-125 per-species scale / median / rate, one entry per observable
    assert abs(_log_prior(error.HalfNormal(), sig) - stats.halfnorm(scale=HINT).logpdf(sig).sum()) < 1e-12
    assert abs(_log_prior(error.LogNormal(sigma=0.7), sig)
               - stats.lognorm(s=0.7, scale=HINT).logpdf(sig).sum()) < 1e-12
    assert abs(_log_prior(error.Gamma(concentration=2.0), sig)
               - stats.gamma(a=2.0, scale=HINT / 2.0).logpdf(sig).sum()) < 1e-12
    # distinct hints give distinct scales: swapping two hints changes the density
    assert abs(_log_prior(error.HalfNormal(), sig) - _log_prior(error.HalfNormal(), sig, HINT[::-1])) > 1.0
    # :131-143 shared=True: ONE scalar site whose scale is the mean of the hint
    s1 = np.array([0.9])
    assert abs(_log_prior(error.HalfNormal(shared=True), s1)
               - stats.halfnorm(scale=HINT.mean()).logpdf(0.9)) < 1e-12
    assert abs(_log_prior(error.LogNormal(shared=True), s1)
               - stats.lognorm(s=0.7, scale=HINT.mean()).logpdf(0.9)) < 1e-12
    # :147-159 an explicit scale / rate ignores the hint
    assert abs(_log_prior(error.HalfNormal(scale=0.3), sig) - stats.halfnorm(scale=0.3).logpdf(sig).sum()) < 1e-12
    assert abs(_log_prior(error.Gamma(concentration=2.0, rate=4.0), sig)
               - stats.gamma(a=2.0, scale=0.25).logpdf(sig).sum()) < 1e-12


def test_bayesian_model_uses_the_per_species_hint_and_the_shared_scalar_site():
    """mcmc.py:464-465 (scale hint from yerrs) and integration/test_mcmc.py:241-256 / 258-293
    (shared collapses to a scalar; per-species yerrs arrays) through ``log_density`` with a
    zero likelihood block standing in for the kernel."""
    import catalax_b200 as ctx
    from catalax_b200.mcmc import BayesianModel
    from scipy import stats

    ctx.enable_x64(True)
    try:
        model = ctx.Model(name="two")
        model.add_state(a="a", b="b")
        model.add_ode("a", "-k * a"); model.add_ode("b", "k * a")
        model.parameters.k.prior = priors.Uniform(0.1, 2.0)
        M, T = 2, 4
        arrays = (np.ones((M, T, 2)), np.tile(np.arange(T, dtype=float), (M, 1)), np.ones((M, 2)),
                  np.zeros((M, 0)))
        theta = torch.full((3, 1), 0.5, dtype=torch.float64)
        zero = lambda th, sg: torch.zeros((th.shape[0], 1 + 1 + 2), dtype=torch.float64)
        log_k = -math.log(1.9)

        bm = BayesianModel(model, arrays=arrays, yerrs=np.array([0.2, 3.0]), shard="none")
        bm._local_loglik = zero
        np.testing.assert_allclose(bm.yerrs, [0.2, 3.0])
        sg = torch.tensor([[0.1, 2.0]] * 3, dtype=torch.float64)
        lp, _, g_sg = bm.log_density(theta, sg)
        want = stats.halfnorm(scale=[0.2, 3.0]).logpdf([0.1, 2.0]).sum() + log_k
        np.testing.assert_allclose(lp.numpy(), want, rtol=1e-13)
        np.testing.assert_allclose(g_sg.numpy(), np.tile([-0.1 / 0.04, -2.0 / 9.0], (3, 1)), rtol=1e-13)

        full = np.broadcast_to(np.array([0.2, 3.0]), (M, T, 2)) * np.linspace(0.5, 1.5, T)[None, :, None]
        assert np.allclose(BayesianModel(model, arrays=arrays, yerrs=full, shard="none").yerrs, [0.2, 3.0])

        sh = BayesianModel(model, arrays=arrays, yerrs=np.array([0.2, 3.0]), shard="none",
                           error_model=error.HalfNormal(shared=True))
        sh._local_loglik = lambda th, sg: torch.cat([torch.zeros((th.shape[0], 2), dtype=torch.float64),
                                                     sg * 10.0], dim=1)     # d/dsigma_j = 10 sigma_j
        lp, _, g_sg = sh.log_density(theta, torch.full((3, 1), 0.7, dtype=torch.float64))
        np.testing.assert_allclose(lp.numpy(), stats.halfnorm(scale=1.6).logpdf(0.7) + log_k, rtol=1e-13)
        assert g_sg.shape == (3, 1)                       # sum of the two partials + the prior's
        np.testing.assert_allclose(g_sg.numpy(), 2 * 7.0 - 0.7 / 1.6 ** 2, rtol=1e-13)
    finally:
        ctx.enable_x64(False)
