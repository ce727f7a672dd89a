"""``HMC.run`` / ``run_mcmc`` entry points (reference: catalax/mcmc/mcmc.py:109-310, 494-538).

The reference hands its numpyro model to NUTS; numpyro is not part of this engine (SURVEY §2:
the sampler is out of scope, the HOT PATH is the log-density + gradient it calls once per leapfrog
step).  So that the reference's entry points work end to end, this module drives the CUDA
log-density (``BayesianModel.log_density``: one fused kernel launch evaluates ALL chains) with a
compact batched Hamiltonian Monte Carlo sampler:

* unconstrained parametrisation as numpyro's ``biject_to``: bounded supports (Uniform,
  LogUniform, TruncatedNormal) through ``low + (high - low) * sigmoid(u)``, positive noise scales
  through ``exp(u)``, with the log-Jacobians added to the density;
* leapfrog integration with a jittered number of steps, per-chain step size by dual averaging
  towards ``target_accept_prob`` and a diagonal mass matrix from the warm-up draws.

It is NOT numpyro's NUTS (no tree doubling): ``max_tree_depth`` only bounds the number of
leapfrog steps per iteration (``2**depth - 1``, capped at 31).  Any other sampler can be plugged in
on ``results.bayesian_model.log_density``.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from . import priors as _priors
from .mcmc import BayesianModel, MCMCConfig


class HMCResults:
    """Posterior draws + the model they belong to (results.py:53-: ``mcmc``, ``bayesian_model``,
    ``model``; ``get_samples`` as numpyro's ``MCMC.get_samples``)."""

    def __init__(self, samples: Dict[str, np.ndarray], bayesian_model, model, stats=None):
        self._samples = samples              # name -> [n_chain, n_draw, ...]
        self.bayesian_model = bayesian_model
        self.model = model
        self.stats = stats or {}
        self.mcmc = self                     # results.mcmc.get_samples(...) as in the reference

    def get_samples(self, group_by_chain: bool = False) -> Dict[str, np.ndarray]:
        if group_by_chain:
            return dict(self._samples)
        return {k: v.reshape(v.shape[0] * v.shape[1], *v.shape[2:]) for k, v in self._samples.items()}

    def summary(self) -> Dict[str, Dict[str, float]]:
        out = {}
        for k, v in self.get_samples().items():
            flat = v.reshape(v.shape[0], -1)
            for j in range(flat.shape[1]):
                name = k if flat.shape[1] == 1 else f"{k}[{j}]"
                col = flat[:, j]
                out[name] = {"mean": float(col.mean()), "sd": float(col.std(ddof=1)),
                             "hdi_3%": float(np.quantile(col, 0.03)), "hdi_97%": float(np.quantile(col, 0.97))}
        return out

    def print_summary(self):
        for name, row in self.summary().items():
            print(f"{name:>12s}  mean {row['mean']:.5g}  sd {row['sd']:.3g}  "
                  f"[{row['hdi_3%']:.5g}, {row['hdi_97%']:.5g}]")


class _Transform:
    """Constrained <-> unconstrained map of one scalar site."""

    def __init__(self, low=None, high=None):
        self.low, self.high = low, high

    def forward(self, u):
        import torch

        if self.low is None:
            return u, torch.zeros_like(u), torch.ones_like(u), torch.zeros_like(u)
        if self.high is None:            # positive: x = low + exp(u)
            e = torch.exp(u)
            return self.low + e, u, e, torch.ones_like(u)
        s = torch.sigmoid(u)
        w = self.high - self.low
        ljac = math.log(w) + torch.nn.functional.logsigmoid(u) + torch.nn.functional.logsigmoid(-u)
        return self.low + w * s, ljac, w * s * (1 - s), 1 - 2 * s     # x, log|J|, dx/du, dlog|J|/du

    def inverse(self, x):
        import torch

        if self.low is None:
            return x
        if self.high is None:
            return torch.log(x - self.low)
        p = ((x - self.low) / (self.high - self.low)).clamp(1e-6, 1 - 1e-6)
        return torch.log(p) - torch.log1p(-p)


def _transform_of(prior) -> _Transform:
    if isinstance(prior, (_priors.Uniform, _priors.LogUniform, _priors.TruncatedNormal)):
        return _Transform(float(prior.low), float(prior.high))
    return _Transform()


def hmc_sample(logp_and_grad, u, lp, g, *, num_warmup, num_samples, thinning=1, max_steps=31,
               target_accept=0.8, generator=None, verbose=False):
    """Batched HMC over ``CH`` chains in unconstrained space.  ``logp_and_grad(u [CH,dim]) ->
    (logp [CH], grad [CH,dim])`` is evaluated for ALL chains at once (one fused kernel launch per
    leapfrog step on the GPU).  Per-chain step size by dual averaging (Hoffman & Gelman 2014,
    restarted after every mass-matrix update), diagonal mass matrix from the warm-up draws (three
    windows, Stan's shrinkage), 1..max_steps leapfrog steps per iteration (jittered, shared by the
    chains so that they stay in lock step).  Returns (list of kept states [CH,dim], stats)."""
    import torch

    CH, dim = u.shape
    dev, tdt = u.device, u.dtype
    gen = generator
    rand = lambda *shape: torch.rand(shape, generator=gen, device=dev, dtype=tdt)
    randn = lambda *shape: torch.randn(shape, generator=gen, device=dev, dtype=tdt)
    inv_mass = torch.ones((CH, dim), dtype=tdt, device=dev)
    eps = torch.full((CH,), 0.1, dtype=tdt, device=dev)
    mu_da = torch.log(10 * eps)
    h_bar = torch.zeros_like(eps); log_eps_bar = torch.log(eps)
    gamma, t0_da, kappa, m_da = 0.05, 10.0, 0.75, 0
    keep, accs, n_div, warm_buf = [], [], 0, []
    windows = sorted({int(num_warmup * f) for f in (0.3, 0.55, 0.8)} - {0})
    for it in range(num_warmup + num_samples):
        L = int(torch.randint(1, max_steps + 1, (1,), generator=gen, device=dev).item())
        r = randn(CH, dim) / torch.sqrt(inv_mass)
        h0 = -lp + 0.5 * (r * r * inv_mass).sum(1)
        un, rn, lpn, gn = u, r, lp, g
        e = eps[:, None]
        for _ in range(L):
            rn = rn + 0.5 * e * gn
            un = un + e * inv_mass * rn
            lpn, gn = logp_and_grad(un)
            rn = rn + 0.5 * e * gn
        h1 = -lpn + 0.5 * (rn * rn * inv_mass).sum(1)
        dH = h0 - h1
        dH = torch.where(torch.isfinite(dH), dH, torch.full_like(dH, -math.inf))
        acc_p = torch.exp(dH.clamp(max=0.0))
        take = rand(CH) < acc_p
        u = torch.where(take[:, None], un, u)
        lp = torch.where(take, lpn, lp)
        g = torch.where(take[:, None], gn, g)
        if it < num_warmup:
            m_da += 1
            h_bar = (1 - 1 / (m_da + t0_da)) * h_bar + (target_accept - acc_p) / (m_da + t0_da)
            log_eps = mu_da - math.sqrt(m_da) / gamma * h_bar
            eta = m_da ** (-kappa)
            log_eps_bar = eta * log_eps + (1 - eta) * log_eps_bar
            eps = torch.exp(log_eps)
            warm_buf.append(u.clone())
            if (it + 1) in windows and len(warm_buf) >= 20:
                w = torch.stack(warm_buf[len(warm_buf) // 2:], 0)            # [n, CH, dim]
                n_w = w.shape[0]
                inv_mass = (n_w / (n_w + 5.0)) * w.var(0, unbiased=True) + 1e-3 * (5.0 / (n_w + 5.0))
                warm_buf = []
                # restart the step-size search around the averaged value
                eps = torch.exp(log_eps_bar)
                mu_da = torch.log(10 * eps); h_bar = torch.zeros_like(eps); m_da = 0
            if it == num_warmup - 1:
                eps = torch.exp(log_eps_bar)
        else:
            accs.append(float(acc_p.mean().item()))
            n_div += int((dH < -1000).sum().item())
            if (it - num_warmup) % thinning == 0:
                keep.append(u.clone())
        if verbose and (it + 1) % max(1, (num_warmup + num_samples) // 10) == 0:
            print(f"sample: {it + 1}/{num_warmup + num_samples}  step size {eps.mean().item():.3g}  "
                  f"acc. prob {acc_p.mean().item():.2f}", flush=True)
    stats = {"step_size": eps.cpu().numpy(), "mean_accept_prob": float(np.mean(accs)) if accs else None,
             "num_divergences": n_div, "inverse_mass": inv_mass.cpu().numpy(),
             "sampler": "batched HMC (dual averaging, diagonal mass); not numpyro NUTS"}
    return keep, stats


class HMC:
    """Same constructor as the reference's ``HMC`` (mcmc.py:116-168)."""

    def __init__(self, num_warmup: int, num_samples: int, likelihood="softlaplace",
                 dense_mass: bool = True, thinning: int = 1, max_tree_depth: int = 10,
                 dt0: float = 0.1, chain_method: str = "sequential", num_chains: int = 1,
                 seed: int = 420, verbose: int = 1, max_steps: int = 64 ** 4, solver=None,
                 target_accept_prob: float = 0.8, error_model=None):
        from ..solvers import Tsit5

        self.config = MCMCConfig(num_warmup=num_warmup, num_samples=num_samples, likelihood=likelihood,
                                 dense_mass=dense_mass, thinning=thinning, max_tree_depth=max_tree_depth,
                                 dt0=dt0, solver=solver or Tsit5, chain_method=chain_method,
                                 num_chains=num_chains, seed=seed, verbose=verbose, max_steps=max_steps,
                                 target_accept_prob=target_accept_prob, error_model=error_model)
        self.likelihood = likelihood

    @classmethod
    def from_config(cls, config: MCMCConfig) -> "HMC":
        inst = cls.__new__(cls)
        inst.config = config
        inst.likelihood = config.likelihood
        return inst

    def run(self, model, dataset, yerrs, surrogate=None, pre_model=None, post_model=None,
            abort_on_error: bool = False, device=None) -> HMCResults:
        """Infer the posterior of the model's parameters and noise scales (mcmc.py:170-293)."""
        import torch

        if pre_model is not None or post_model is not None:
            raise NotImplementedError("pre_model / post_model hooks are jax callables in the reference")
        missing = [n for n in model.get_parameter_order() if model.parameters[n].prior is None]
        assert not missing, f"Parameters {missing} do not have priors. Please specify priors for all parameters."
        cfg = self.config
        bm = BayesianModel(model, dataset, yerrs=yerrs, likelihood=cfg.likelihood,
                           config=cfg.to_simulation_config(), error_model=cfg.error_model,
                           surrogate=surrogate, shard="none")
        dev = device or torch.device("cuda", torch.cuda.current_device())
        tdt = torch.float64 if np.dtype(bm.dtype) == np.float64 else torch.float32
        CH, P = int(cfg.num_chains), bm.field.P
        shared = bool(getattr(bm.error_model, "shared", False))
        sampled = bool(bm.error_model.sampled)
        n_sig = (1 if shared else bm.field.S)
        tf_theta = [_transform_of(pr) for _, pr in bm.priors]
        tf_sig = _Transform(0.0, None)
        gen = torch.Generator(device=dev).manual_seed(int(cfg.seed))
        fixed_sigma = None
        if not sampled:
            fixed_sigma = torch.as_tensor(np.broadcast_to(np.asarray(bm.error_model.sigma_y, dtype=np.float64),
                                                          (n_sig,)).copy(), dtype=tdt, device=dev)
        dim = P + (n_sig if sampled else 0)

        def constrain(u):
            xs, lj, dx, dlj = [], [], [], []
            for j in range(P):
                a, b, c, d = tf_theta[j].forward(u[:, j])
                xs.append(a); lj.append(b); dx.append(c); dlj.append(d)
            for j in range(P, dim):
                a, b, c, d = tf_sig.forward(u[:, j])
                xs.append(a); lj.append(b); dx.append(c); dlj.append(d)
            st = lambda v: torch.stack(v, 1)
            return st(xs), st(lj).sum(1), st(dx), st(dlj)

        def logp_and_grad(u):
            x, ljac, dx, dlj = constrain(u)
            theta = x[:, :P].contiguous()
            sigma = x[:, P:].contiguous() if sampled else fixed_sigma.expand(u.shape[0], -1).contiguous()
            lp, g_th, g_sg = bm.log_density(theta, sigma)
            g = torch.cat([g_th, g_sg], 1) if sampled else g_th
            lp = lp + ljac
            g = g * dx + dlj
            bad = ~torch.isfinite(lp) | ~torch.isfinite(g).all(1)
            lp = torch.where(bad, torch.full_like(lp, -math.inf), lp)
            g = torch.where(bad[:, None], torch.zeros_like(g), g)
            return lp, g

        # initial point: numpyro's init_to_uniform(radius=2) in unconstrained space, retried where
        # the density is not finite
        u = (torch.rand((CH, dim), generator=gen, device=dev, dtype=tdt) * 4 - 2)
        lp, g = logp_and_grad(u)
        for _ in range(50):
            bad = ~torch.isfinite(lp)
            if not bool(bad.any()):
                break
            u = torch.where(bad[:, None], torch.rand((CH, dim), generator=gen, device=dev, dtype=tdt) * 4 - 2, u)
            lp, g = logp_and_grad(u)
        if bool((~torch.isfinite(lp)).any()):
            raise RuntimeError("could not find an initial point with a finite log-density")

        keep_u, stats = hmc_sample(logp_and_grad, u, lp, g, num_warmup=int(cfg.num_warmup),
                                   num_samples=int(cfg.num_samples), thinning=max(1, int(cfg.thinning)),
                                   max_steps=min(2 ** int(cfg.max_tree_depth) - 1, 31),
                                   target_accept=float(cfg.target_accept_prob), generator=gen,
                                   verbose=bool(cfg.verbose))
        keep = [constrain(uu)[0].cpu().numpy() for uu in keep_u]
        draws = np.stack(keep, 1) if keep else np.zeros((CH, 0, dim))     # [CH, n_draw, dim]
        samples = {name: draws[:, :, j] for j, (name, _) in enumerate(bm.priors)}
        if sampled:
            samples["sigma_y"] = draws[:, :, P:]
        return HMCResults(samples, bm, model, stats)


def run_mcmc(model, dataset, yerrs, config: MCMCConfig, surrogate=None, pre_model=None,
             post_model=None) -> HMCResults:
    """mcmc.py:494-538."""
    return HMC.from_config(config).run(model=model, dataset=dataset, yerrs=yerrs, surrogate=surrogate,
                                       pre_model=pre_model, post_model=post_model)
