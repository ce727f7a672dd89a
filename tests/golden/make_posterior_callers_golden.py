"""Golden arrays for the posterior re-integration callers, produced by executing the
reference's own functions (run in the build container, where /root/reference exists).

Executed UNMODIFIED, extracted with ``ast`` from catalax/mcmc/loo.py and catalax/mcmc/predictive.py:

  loo.py         _stack_thetas :92, _sigma_y_draws :103, _subsample :120, _mechanistic_loglik :131,
                 _surrogate_loglik :203, _resolve_yerrs :306, reconstruct_log_likelihood :331 (point /
                 curve leave-out, both fit modes)
  predictive.py  euler_integrate_rates :54, _integrate_ensemble :109, _aleatoric_variance :158

over stand-ins: ``jax.numpy`` -> numpy, ``jax.vmap(f)`` -> a Python loop over the leading axis,
``catalax.mcmc.mcmc._prepare_mcmc_data`` -> a record holding the arrays and the parity oracle's solve
as ``sim_func``, ``numpyro`` distributions -> the restated log-densities of
make_logjoint_golden.py, ``results`` -> a record with the attributes those functions read.
Output: tests/golden/posterior_callers_reference_lines.npz
"""
import ast
import sys
import types
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_logjoint_golden as lj   # noqa: E402  (the numpyro / jax.numpy stand-ins)

REF = Path("/root/reference/catalax/mcmc")
OUT = Path(__file__).parent / "posterior_callers_reference_lines.npz"


def definitions(path, *names):
    text = path.read_text()
    lines = text.splitlines()
    parts = []
    for node in ast.parse(text).body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            parts.append("\n".join(lines[node.lineno - 1: node.end_lineno]))
    assert len(parts) == len(names), (path, names)
    return "from __future__ import annotations\n" + "\n\n\n".join(parts)


def vmap(f):
    def run(*arrays):
        outs = [f(*(a[i] for a in arrays)) for i in range(arrays[0].shape[0])]
        return np.stack(outs)
    return run


jnp = lj.jnp
for n in ("where", "linspace", "min", "max", "diff", "zeros"):
    setattr(jnp, n, getattr(np, n))
jax = types.SimpleNamespace(vmap=vmap, numpy=jnp)


class SimulationConfig:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def main():
    from catalax_b200 import workloads as wl
    from oracle import diffrax_oracle as do, symbolic as sy

    rng = np.random.default_rng(99)
    mm = wl.michaelis_menten()
    S, P = 3, 2
    M, T, n_chain, n_draw = 5, 9, 3, 8
    rhs, *_ = sy.make_rhs(*mm[:4], mm[4])
    y0s = np.stack([rng.uniform(5, 20, M), np.zeros(M), rng.uniform(50, 300, M)], 1)
    times = np.tile(np.linspace(0, 80, T), (M, 1)) + rng.uniform(0, 2, (M, 1)) * np.arange(T)[None, :] / T
    consts = np.zeros((M, 0))
    clean = do.solve(rhs, y0s, np.array(wl.NUTS_THETA_STAR), consts, times, rtol=1e-8, atol=1e-8).ys
    data = clean + rng.normal(size=clean.shape) * (0.05 * np.abs(clean) + 1e-3)
    data[rng.uniform(size=data.shape) < 0.1] = np.nan
    data[3] = np.nan                                   # a series without any observation
    posterior = {"K_m": wl.NUTS_THETA_STAR[0] * np.exp(rng.normal(0, 0.1, (n_chain, n_draw))),
                 "k_cat": wl.NUTS_THETA_STAR[1] * np.exp(rng.normal(0, 0.1, (n_chain, n_draw))),
                 "sigma_y": np.exp(rng.normal(np.log(0.6), 0.2, (n_chain, n_draw, S)))}
    solve_kw = dict(rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096)

    def sim_func(y0s_, theta_, constants_, times_):
        return do.solve(rhs, y0s_, np.asarray(theta_, dtype=float), constants_, times_, **solve_kw).ys

    Modes = lj.Modes
    prep = types.SimpleNamespace(data=data, mask=np.isfinite(data), sim_func=sim_func, y0s=y0s,
                                 constants=consts, times=times)
    fake_mcmc = types.ModuleType("catalax.mcmc.mcmc")
    fake_mcmc._prepare_mcmc_data = lambda dataset, model, surrogate=None, config=None: prep
    fake_mcmc.Modes = Modes
    fake_diffrax = types.ModuleType("diffrax")
    fake_diffrax.Tsit5 = "tsit5"
    sys.modules.update({"catalax": types.ModuleType("catalax"), "catalax.mcmc": types.ModuleType("catalax.mcmc"),
                        "catalax.mcmc.mcmc": fake_mcmc, "diffrax": fake_diffrax})

    ns = {"np": np, "jnp": jnp, "jax": jax, "SimulationConfig": SimulationConfig}
    exec(compile(definitions(REF / "loo.py", "_build_eval_config", "_stack_thetas", "_sigma_y_draws",
                             "_subsample", "_mechanistic_loglik", "_resolve_yerrs",
                             "reconstruct_log_likelihood"), str(REF / "loo.py"), "exec"), ns)
    ns_p = dict(ns)
    ns_p["_stack_thetas"] = ns["_stack_thetas"]
    exec(compile(definitions(REF / "predictive.py", "_integrate_ensemble", "_aleatoric_variance"),
                 str(REF / "predictive.py"), "exec"), ns_p)

    out = dict(y0s=y0s, times=times, data=data, post_K_m=posterior["K_m"], post_k_cat=posterior["k_cat"],
               post_sigma_y=posterior["sigma_y"])
    for lik in ("Normal", "SoftLaplace"):
        bm = types.SimpleNamespace(observables=np.arange(S), likelihood=getattr(lj.dist, lik),
                                   priors=[("K_m", None), ("k_cat", None)], yerrs=np.broadcast_to(0.8, data.shape),
                                   mode=Modes.MECHANISTIC, dt0=0.1, solver="tsit5")
        model = types.SimpleNamespace(get_observable_state_order=lambda: list(mm[0]))
        results = types.SimpleNamespace(
            bayesian_model=bm, model=model,
            mcmc=types.SimpleNamespace(get_samples=lambda group_by_chain=True: posterior))
        tag = lik.lower()
        for leave_out in ("point", "curve"):
            for max_draws in (None, 3):
                ll, post_np, meta = ns["reconstruct_log_likelihood"](
                    results, None, leave_out=leave_out, max_draws=max_draws)
                key = f"{tag}/{leave_out}/{max_draws}"
                out[f"{key}/loglik"] = ll
                out[f"{key}/kept_idx"] = meta["kept_idx"]
                out[f"{key}/n_dropped"] = np.array(meta["n_dropped"])
                out[f"{key}/post_K_m"] = post_np["K_m"]
        ll_y, _, meta = ns["reconstruct_log_likelihood"](results, None, sigma_source="yerrs",
                                                         yerrs=np.array([0.5, 1.0, 2.0]), max_draws=3)
        out[f"{tag}/yerrs/loglik"] = ll_y
        full, mask, info = ns["_mechanistic_loglik"](results, None, posterior, sigma_source="reuse",
                                                     yerrs=None, config=None)
        out[f"{tag}/full"], out["mask"] = full, mask
    # ---- surrogate-mode fit: loo.py:203-303 (_surrogate_loglik) + predictive.py:54-106
    # (euler_integrate_rates): the sampled mechanistic rates at the measured states, forward-Euler
    # integrated along the measurement times and scored against the measured concentrations
    full = np.nan_to_num(clean) * np.exp(rng.normal(0, 0.02, clean.shape))
    full[2, 4, 1] = np.nan                              # a missing concentration
    exec(compile(definitions(REF / "predictive.py", "euler_integrate_rates"), str(REF / "predictive.py"), "exec"), ns_p)
    fake_pred = types.ModuleType("catalax.mcmc.predictive")
    fake_pred.euler_integrate_rates = ns_p["euler_integrate_rates"]
    sys.modules["catalax.mcmc.predictive"] = fake_pred
    for n in ("cumsum", "concatenate"):
        setattr(jnp, n, getattr(np, n))
    exec(compile(definitions(REF / "loo.py", "_surrogate_loglik"), str(REF / "loo.py"), "exec"), ns)

    er = dict(rates=rng.normal(size=(M, T, S)), y0_obs=rng.normal(size=(M, S)), dt=rng.uniform(0.1, 1.0, (M, T - 1)),
              rate_var=rng.uniform(0, 1, (M, T, S)), data_safe=rng.normal(size=(M, T, S)))
    for k, v in er.items():
        out[f"euler/{k}"] = v
    for mode in ("euler", "euler_onestep"):
        yh, vy = ns_p["euler_integrate_rates"](er["rates"], er["y0_obs"], er["dt"], er["rate_var"],
                                               integration=mode, data_safe=er["data_safe"])
        out[f"euler/{mode}/yhat"], out[f"euler/{mode}/var_y"] = yh, vy

    def rate_sim_func(points, theta_, constants_, times_):
        n = points.shape[0]
        ini = np.repeat(points.reshape(-1, T, S)[:, 0, :], T, axis=0)
        return rhs(times_, points, np.broadcast_to(np.asarray(theta_, dtype=float), (n, P)), constants_, ini)

    dataset = types.SimpleNamespace(to_jax_arrays=lambda order: (full, times, None),
                                    to_y0_matrix=lambda state_order: consts)
    out["surr/full_data"] = full
    for lik in ("Normal", "SoftLaplace"):
        bm = types.SimpleNamespace(observables=np.arange(S), likelihood=getattr(lj.dist, lik),
                                   priors=[("K_m", None), ("k_cat", None)], sim_func=rate_sim_func,
                                   yerrs=np.broadcast_to(0.8, (M * T, S)), mode=Modes.SURROGATE, dt0=0.1,
                                   solver="tsit5")
        smodel = types.SimpleNamespace(get_state_order=lambda modeled=False: list(mm[0]),
                                       get_constants_order=lambda: [],
                                       get_observable_state_order=lambda: list(mm[0]))
        sresults = types.SimpleNamespace(
            bayesian_model=bm, model=smodel,
            mcmc=types.SimpleNamespace(get_samples=lambda group_by_chain=True: posterior))
        for integration in ("euler", "euler_onestep"):
            ll, post_np, meta = ns["reconstruct_log_likelihood"](sresults, dataset, integration=integration,
                                                                 max_draws=4)
            out[f"surr/{lik.lower()}/{integration}/loglik"] = ll
            out[f"surr/{lik.lower()}/{integration}/kept_idx"] = meta["kept_idx"]
        ll, _, _ = ns["reconstruct_log_likelihood"](sresults, dataset, sigma_source="yerrs", yerrs=1.3,
                                                    leave_out="curve", max_draws=4)
        out[f"surr/{lik.lower()}/yerrs_curve/loglik"] = ll

    # ---- predictive: dense per-series grids, ensemble, aleatoric variance
    sub = ns["_subsample"](posterior, 4)
    t_dense, yhat = ns_p["_integrate_ensemble"](results, None, sub, n_steps=17)
    out["ens/times"], out["ens/yhat"] = t_dense, yhat
    out["ens/var"] = ns_p["_aleatoric_variance"](results, sub, yhat.shape)
    out["ens/sub_K_m"] = sub["K_m"]
    # ---- bands: predictive.py:252-293 (band_from_ensemble) with its _PERCENTILE table; the Dataset
    # wrapping is replaced by a pass-through, jax.random.normal(PRNGKey(0), shape) by the Threefry
    # restatement of catalax_b200/dataset/jaxrandom.py (pinned on jax's known answers separately)
    from catalax_b200.dataset.jaxrandom import normal as threefry_normal

    ptext = (REF / "predictive.py").read_text()
    for node in ast.parse(ptext).body:
        if isinstance(node, ast.AnnAssign) and getattr(node.target, "id", "") == "_PERCENTILE":
            ns_p["_PERCENTILE"] = ast.literal_eval(node.value)
    ns_p["_ensemble_to_dataset"] = lambda reference, results_, times_, values: np.asarray(values)
    ns_p["jax"] = types.SimpleNamespace(
        vmap=vmap, numpy=jnp,
        random=types.SimpleNamespace(PRNGKey=lambda seed: seed,
                                     normal=lambda key, shape: threefry_normal(key, int(np.prod(shape)), np.float32).reshape(shape)))
    exec(compile(definitions(REF / "predictive.py", "band_from_ensemble"), str(REF / "predictive.py"), "exec"), ns_p)
    for hdi in (None, "lower", "upper", "lower_50", "upper_50"):
        for noise in (True, False):
            out[f"band/{hdi}/{noise}"] = ns_p["band_from_ensemble"](results, None, t_dense, yhat, out["ens/var"],
                                                                   hdi=hdi, include_noise=noise)
    np.savez_compressed(OUT, **out)
    for k in sorted(out):
        print(k, out[k].shape)


if __name__ == "__main__":
    main()
