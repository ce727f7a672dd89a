"""Golden values of the NUTS log-joint, produced by executing the reference's own source (run in
the build container, where /root/reference exists; the fixture travels).

numpyro / jax cannot be imported here, so ``run_mcmc`` cannot run.  What CAN run is the text that
DEFINES the density:

  * catalax/mcmc/priors.py   -- the whole file (pydantic models + ``_distribution_fun``)
  * catalax/mcmc/error.py    -- the whole file (HalfNormal / LogNormal / Gamma / Fixed ``sample``,
                                 ``_resolve_hint``, ``shared``)
  * catalax/mcmc/mcmc.py     -- ``Modes``, ``_per_species_scale_hint`` and the complete class
                                 ``BayesianModel`` (``__init__`` :340-403 with the observable block of
                                 the rate-Jacobian, ``__call__`` :405-491 with the sigma push-forward,
                                 the 1e-8 floor, ``sqrt(sigma_total)``, the mask)

They are executed UNMODIFIED against stand-ins for the two libraries they call:

  * ``jax.numpy`` -> numpy (array, asarray, einsum, sqrt, broadcast_to, mean, ones, log);
  * ``numpyro``   -> a tracer: ``sample(name, fn, obs=None)`` takes the site value from the trace
    (or ``obs``), records ``fn.log_prob(value)`` summed; ``handlers.mask(mask=m)`` is numpyro's
    MaskedDistribution rule (masked values are replaced by a feasible point, their log-prob by 0);
    ``deterministic`` records nothing;
  * ``numpyro.distributions`` -> Normal, HalfNormal, LogNormal, Gamma, Uniform, LogUniform,
    TruncatedNormal, SoftLaplace with the log-densities numpyro 0.19 documents (RESTATED, the only
    restated part besides the ODE solve).

``sim_func`` is the parity oracle's solve (integrated mode) / rate function (surrogate mode); the
states it returned are stored next to the results, so the fixture pins everything the reference
does AROUND the solve: parameter order, prior objects and defaults, the error model's scale
hint, shared / per-species noise, the push-forward, the masked likelihood.
Output: tests/golden/logjoint_reference_lines.npz
"""
import ast
import contextlib
import importlib.util
import math
import sys
import types
from copy import deepcopy
from enum import Enum, auto
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
REF = Path("/root/reference/catalax/mcmc")
OUT = Path(__file__).parent / "logjoint_reference_lines.npz"


# ------------------------------------------------------------------ numpyro.distributions stand-in
def _phi(x):
    return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))


class Distribution:
    pass


class Normal(Distribution):
    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = np.asarray(loc, dtype=float), np.asarray(scale, dtype=float)

    def log_prob(self, x):
        z = (x - self.loc) / self.scale
        return -0.5 * z * z - np.log(self.scale) - 0.5 * math.log(2 * math.pi)

    feasible = 0.0


class SoftLaplace(Normal):
    def log_prob(self, x):
        z = (x - self.loc) / self.scale
        return math.log(2 / math.pi) - np.log(self.scale) - np.logaddexp(z, -z)


class HalfNormal(Distribution):
    def __init__(self, scale=1.0):
        self.scale = np.asarray(scale, dtype=float)

    def log_prob(self, x):
        return Normal(0.0, self.scale).log_prob(x) + math.log(2.0)


class LogNormal(Distribution):
    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = np.asarray(loc, dtype=float), np.asarray(scale, dtype=float)

    def log_prob(self, x):
        return Normal(self.loc, self.scale).log_prob(np.log(x)) - np.log(x)


class Gamma(Distribution):
    def __init__(self, concentration, rate=1.0):
        self.a, self.rate = np.asarray(concentration, dtype=float), np.asarray(rate, dtype=float)

    def log_prob(self, x):
        lg = np.vectorize(math.lgamma)(self.a)
        return self.a * np.log(self.rate) + (self.a - 1) * np.log(x) - self.rate * x - lg


class Uniform(Distribution):
    def __init__(self, low=0.0, high=1.0):
        self.low, self.high = float(low), float(high)

    def log_prob(self, x):
        return np.where((x >= self.low) & (x <= self.high), -math.log(self.high - self.low), -np.inf)


class LogUniform(Distribution):
    def __init__(self, low, high):
        self.low, self.high = float(low), float(high)

    def log_prob(self, x):      # exp-transform of Uniform(log low, log high)
        return -np.log(x) - math.log(math.log(self.high) - math.log(self.low))


class TruncatedNormal(Distribution):
    def __init__(self, loc=0.0, scale=1.0, *, low=None, high=None):
        self.loc, self.scale, self.low, self.high = float(loc), float(scale), float(low), float(high)

    def log_prob(self, x):
        z = _phi((self.high - self.loc) / self.scale) - _phi((self.low - self.loc) / self.scale)
        return Normal(self.loc, self.scale).log_prob(x) - math.log(z)


dist = types.ModuleType("numpyro.distributions")
for _c in (Distribution, Normal, SoftLaplace, HalfNormal, LogNormal, Gamma, Uniform, LogUniform,
           TruncatedNormal):
    setattr(dist, _c.__name__, _c)


# ------------------------------------------------------------------ numpyro stand-in (tracer)
class Trace:
    values: dict = {}
    logp: dict = {}
    mask = None


def sample(name, fn, obs=None):
    value = obs if obs is not None else Trace.values[name]
    value = np.asarray(value, dtype=float)
    if Trace.mask is not None:      # MaskedDistribution.log_prob
        m = np.asarray(Trace.mask, dtype=bool)
        lp = np.where(m, fn.log_prob(np.where(m, value, 0.0)), 0.0)
    else:
        lp = fn.log_prob(value)
    Trace.logp[name] = float(np.sum(lp))
    return value


@contextlib.contextmanager
def _mask(mask=None):
    Trace.mask = mask
    try:
        yield
    finally:
        Trace.mask = None


numpyro = types.ModuleType("numpyro")
numpyro.sample = sample
numpyro.deterministic = lambda name, value: value
numpyro.handlers = types.SimpleNamespace(mask=_mask)
numpyro.distributions = dist

jnp = types.ModuleType("jax.numpy")
for _n in ("array", "asarray", "einsum", "sqrt", "broadcast_to", "mean", "ones", "log", "sum", "repeat"):
    setattr(jnp, _n, getattr(np, _n))
jax = types.ModuleType("jax")
jax.numpy = jnp
jax.Array = np.ndarray


def load_reference_module(name):
    """Execute a whole reference file with the stand-in libraries in sys.modules."""
    fakes = {"jax": jax, "jax.numpy": jnp, "numpyro": numpyro, "numpyro.distributions": dist}
    saved = {k: sys.modules.get(k) for k in fakes}
    sys.modules.update(fakes)
    try:
        spec = importlib.util.spec_from_file_location(f"_ref_{name}", REF / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


ref_priors = load_reference_module("priors")
ref_error = load_reference_module("error")


def reference_definitions(*names):
    """Source text of top-level classes / functions of catalax/mcmc/mcmc.py, unmodified."""
    text = (REF / "mcmc.py").read_text()
    lines = text.splitlines()
    parts = []
    for node in ast.parse(text).body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            start = min([node.lineno] + [d.lineno for d in node.decorator_list])
            parts.append("\n".join(lines[start - 1: node.end_lineno]))
    assert len(parts) == len(names), names
    return "from __future__ import annotations\n" + "\n\n\n".join(parts)


NS = {"jnp": jnp, "numpyro": numpyro, "dist": dist, "deepcopy": deepcopy, "Enum": Enum, "auto": auto,
      "DEFAULT_ERROR_MODEL": ref_error.DEFAULT_ERROR_MODEL}
exec(compile(reference_definitions("Modes", "_per_species_scale_hint", "BayesianModel"),
             str(REF / "mcmc.py"), "exec"), NS)
Modes, RefBayesianModel = NS["Modes"], NS["BayesianModel"]


class StandInModel:
    """The four things BayesianModel.__init__ asks of a catalax Model."""

    def __init__(self, states, param_priors):
        self._states = list(states)
        self.parameters = {n: types.SimpleNamespace(name=n, prior=p) for n, p in param_priors.items()}

    def _setup_system(self, config):
        pass

    def get_parameter_order(self):
        return sorted(self.parameters)

    def get_state_order(self, as_indices=False, modeled=False):
        return list(range(len(self._states))) if as_indices else list(self._states)


def log_joint(bm, draws, args):
    """Run the reference model once per draw; per-site summed log-probabilities."""
    rows = []
    for values in draws:
        Trace.values, Trace.logp = values, {}
        bm(*args)
        rows.append(dict(Trace.logp))
    sites = list(rows[0])
    return {s: np.array([r[s] for r in rows]) for s in sites}


def main():
    from catalax_b200 import workloads as wl
    from oracle import diffrax_oracle as do, symbolic as sy

    rng = np.random.default_rng(20241020)
    model = wl.michaelis_menten()                 # states [E, P, S], parameters [K_m, k_cat]
    states, params = model[0], model[1]
    S, P = len(states), len(params)
    M, T, CH = 6, 12, 5
    config = types.SimpleNamespace(solver="tsit5", dt0=0.1)
    solve_kw = dict(rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096)    # mcmc.py:73-79

    S0 = rng.uniform(50, 300, M); E0 = rng.uniform(5, 20, M)
    y0s = np.stack([E0, np.zeros(M), S0], axis=1)
    times = np.tile(np.linspace(0, 100, T), (M, 1))
    times[-1] = np.linspace(0, 60, T)                                  # per-series grids
    consts = np.zeros((M, 0))
    rhs, *_ = sy.make_rhs(*model[:4], model[4])
    clean = do.solve(rhs, y0s, np.array(wl.NUTS_THETA_STAR), consts, times, rtol=1e-8, atol=1e-8).ys
    data = clean + rng.normal(size=clean.shape) * (0.05 * np.abs(clean) + 1e-3)
    data[rng.uniform(size=data.shape) < 0.08] = np.nan
    data[-1, 9:, :] = np.nan                                           # ragged last series
    mask = np.isfinite(data)                                           # mcmc.py:566
    theta = np.array(wl.NUTS_THETA_STAR) * np.exp(rng.normal(0, 0.2, (CH, P)))
    sig3 = np.exp(rng.normal(np.log(0.5), 0.3, (CH, S)))
    sig1 = np.exp(rng.normal(np.log(0.5), 0.3, (CH, 1)))

    sim_states = {}

    def sim_integrated(y0s_, theta_, constants_, times_):
        ys = do.solve(rhs, y0s_, np.asarray(theta_, dtype=float), constants_, times_, **solve_kw).ys
        sim_states[tuple(np.round(theta_, 12))] = ys
        return ys

    out = dict(y0s=y0s, times=times, data=data, theta=theta, sigma3=sig3, sigma1=sig1,
               states=np.array(states), params=np.array(params))

    prior_sets = {
        "uniform_loguniform": {"K_m": ref_priors.Uniform(low=1.0, high=200.0),
                               "k_cat": ref_priors.LogUniform(low=0.01, high=10.0)},
        "normal_truncnormal": {"K_m": ref_priors.Normal(mu=45.0, sigma=12.0),
                               "k_cat": ref_priors.TruncatedNormal(mu=0.25, sigma=0.2)},
    }
    yerrs_species = np.broadcast_to(np.array([0.4, 1.5, 2.5]), data.shape).copy()
    cases = {
        # name: (priors, likelihood, error model, yerrs, shared?)
        "A": ("uniform_loguniform", "SoftLaplace", None, np.broadcast_to(0.8, data.shape), False),
        "B": ("normal_truncnormal", "Normal", ref_error.LogNormal(sigma=0.7), yerrs_species, False),
        "C": ("uniform_loguniform", "SoftLaplace", ref_error.Gamma(concentration=2.0, shared=True),
              yerrs_species, True),
        "D": ("normal_truncnormal", "Normal", ref_error.Fixed(np.array([0.3, 0.9, 1.1])), yerrs_species, False),
        "E": ("uniform_loguniform", "Normal", ref_error.HalfNormal(scale=1.7), yerrs_species, False),
    }
    for name, (pset, lik, em, yerrs, shared) in cases.items():
        bm = RefBayesianModel(model=StandInModel(states, prior_sets[pset]), yerrs=yerrs,
                              likelihood=getattr(dist, lik), sim_func=sim_integrated, shapes=None,
                              mode=Modes.MECHANISTIC, config=config, error_model=em)
        sig = sig1 if shared else sig3
        draws = [dict(K_m=theta[c, 0], k_cat=theta[c, 1],
                      sigma_y=(sig[c, 0] if shared else sig[c])) for c in range(CH)]
        sites = log_joint(bm, draws, (y0s, consts, times, mask, data))
        for s, v in sites.items():
            out[f"{name}/{s}"] = v
        out[f"{name}/meta"] = np.array([pset, lik.lower(), type(bm.error_model).__name__, str(shared)])
        out[f"{name}/yerrs"] = np.asarray(yerrs, dtype=float)
    out["sim_states"] = np.stack([sim_states[tuple(np.round(theta[c], 12))] for c in range(CH)])

    # ---- surrogate (rate-matching) mode: mcmc.py:469-480, :819-847
    N = M * T
    y_pts = np.nan_to_num(clean).reshape(N, S) * np.exp(rng.normal(0, 0.02, (N, S)))
    t_pts = times.ravel()
    inits = y_pts.reshape(M, T, S)[:, 0, :]
    rates = rhs(t_pts, y_pts, np.broadcast_to(np.array(wl.NUTS_THETA_STAR), (N, P)), np.zeros((N, 0)),
                np.repeat(inits, T, axis=0))
    rates = rates + rng.normal(size=rates.shape) * 0.05
    rates[rng.uniform(size=rates.shape) < 0.06] = np.nan
    smask = np.isfinite(rates)
    jac = rng.normal(size=(N, S, S)) * 0.1
    sigma_surr = rng.uniform(0.0, 0.1, (N, S))
    rate_sigma = rng.uniform(0.0, 0.05, (S,))
    out.update(surr_y=y_pts, surr_t=t_pts, surr_rates=rates, surr_jac=jac, surr_sigma_surr=sigma_surr,
               surr_rate_sigma=rate_sigma)

    def sim_rates(y_, theta_, constants_, times_):
        # rate function at the measured states; "{state}_init" = first state of each measurement
        n = y_.shape[0]
        ini = np.repeat(y_.reshape(-1, T, S)[:, 0, :], T, axis=0)
        return rhs(times_, y_, np.broadcast_to(np.asarray(theta_, dtype=float), (n, P)), constants_, ini)

    surr_cases = {
        "F": ("uniform_loguniform", "SoftLaplace", None, sigma_surr, rate_sigma),
        "G": ("normal_truncnormal", "Normal", ref_error.LogNormal(sigma=0.5), None, None),
    }
    for name, (pset, lik, em, ssurr, rsig) in surr_cases.items():
        bm = RefBayesianModel(model=StandInModel(states, prior_sets[pset]),
                              yerrs=np.broadcast_to(0.8, data.shape), likelihood=getattr(dist, lik),
                              sim_func=sim_rates, shapes=None, mode=Modes.SURROGATE, config=config,
                              sigma_surr=ssurr, rate_sigma=rsig, rate_jac=jac, error_model=em)
        draws = [dict(K_m=theta[c, 0], k_cat=theta[c, 1], sigma_y=sig3[c]) for c in range(CH)]
        sites = log_joint(bm, draws, (y_pts, np.zeros((N, 0)), t_pts, smask, rates))
        for s, v in sites.items():
            out[f"{name}/{s}"] = v
        out[f"{name}/meta"] = np.array([pset, lik.lower(), type(bm.error_model).__name__,
                                        str(ssurr is not None)])
    np.savez_compressed(OUT, **out)
    for k in sorted(out):
        if "/" in k and not k.endswith(("meta", "yerrs")):
            print(k, np.array2string(out[k], precision=6))


if __name__ == "__main__":
    main()
