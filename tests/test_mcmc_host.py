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
