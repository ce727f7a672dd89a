"""Pin the oracle against OUTPUTS OF THE REFERENCE ITSELF that ship inside /root/reference.  CPU.

Two artefacts of real reference runs exist besides the test fixtures:

1. ``examples/datasets/croissant_dataset.zip`` -- 120 trajectories written by
   ``examples/GenerateData.ipynb``: ``Model.simulate`` with the DEFAULT config (Tsit5,
   rtol=atol=1e-5, dt0=0.1, f32) of ``ds1/dt = -v_max s1/(K_m+s1)``, ``K_m=100, v_max=7``, 20
   points on [0, 100] (``tests/golden/croissant_menten.npz``, made by
   ``tests/golden/make_croissant_golden.py``).  At this tolerance the discretisation error
   (up to 2e-2 absolute, inside steps of 40 time units) is 100x the f32 rounding, so
   ``reference - closed form`` is a fingerprint of diffrax's accepted steps and of the Tsit5
   dense interpolant -- the step-sequence-level evidence the reference's own asserts
   (<= 1e-3 against closed forms) cannot give.

   One caveat is physical, not an implementation detail: the step [0.1, 1.1] has a true error
   estimate of ~1e-9 relative, i.e. its f32 value is rounding noise, and the controller turns
   that noise into the next step length (factor anywhere in ~[4, 10]).  No second implementation
   (not even XLA on another CPU) reproduces that number.  The test therefore treats that ONE
   factor per series as a nuisance parameter (``factor_override``), fitted on the 20 points.

2. ``examples/HMC.ipynb`` cell 6 prints the posterior summary of a real numpyro NUTS run
   (1000 draws) on that data set after ``augment(seed=0, sigma=5e-2, multiplicative=True)``:
   ``K_m 102.868 +- 1.043, sigma 2.138 +- 0.041, v_max 7.087 +- 0.035`` with MCSEs
   0.048 / 0.002 / 0.002.  Posterior moments are functionals of the log-density, so integrating
   the ORACLE's log-density on a grid must reproduce them to Monte-Carlo error.  The notebook
   predates the ``sigma_y`` error models: its single ``sigma`` site is the "historical default"
   HalfNormal(yerrs) (``catalax/mcmc/error.py`` HalfNormal docstring) with SoftLaplace likelihood.
"""
from pathlib import Path

import numpy as np
import pytest
from scipy.special import lambertw

from oracle import diffrax_oracle as do, jax_prng, loglik_oracle as lo, symbolic as sy
from oracle.c_oracle import COracle

GOLD = Path(__file__).parent / "golden" / "croissant_menten.npz"
MENTEN = (["s1"], ["K_m", "v_max"], [], {"s1": "-(v_max * s1) / (K_m + s1)"}, [])
THETA_TRUE = np.array([100.0, 7.0])

# examples/HMC.ipynb cell 6 (results.summary()): mean, sd, mcse_mean, mcse_sd
HMC_SUMMARY = {"K_m": (102.868, 1.043, 0.048, 0.034),
               "sigma": (2.138, 0.041, 0.002, 0.001),
               "v_max": (7.087, 0.035, 0.002, 0.001)}


def _golden():
    g = np.load(GOLD)
    return g["data"].astype(np.float64), g["times"].astype(np.float64)


def closed_form(t, s0, v, k):
    return k * np.real(lambertw(s0 / k * np.exp((s0 - v * t) / k)))


# ----------------------------------------------------------------------------- jax PRNG KATs
def test_threefry_and_jax_random_known_answers():
    # Random123 kat_vectors: threefry2x32_20, zero counter, zero key
    a, b = jax_prng.threefry2x32(0, 0, [0], [0])
    assert (int(a[0]), int(b[0])) == (0x6B200159, 0x99BA4EFE)
    # values printed by jax's documentation for key(0) / key(42): partitionable layout (jax >= 0.5)
    assert abs(float(jax_prng.uniform(0, 1)[0]) - 0.947667) < 1e-6
    assert abs(float(jax_prng.normal(0, 1)[0]) - 1.6226422) < 2e-6
    assert abs(float(jax_prng.normal(42, 1)[0]) + 0.028304616) < 1e-6
    # ... and the legacy layout (jax < 0.5)
    assert abs(float(jax_prng.uniform(0, 1, partitionable=False)[0]) - 0.41845703) < 1e-7
    assert abs(float(jax_prng.normal(0, 1, partitionable=False)[0]) + 0.20584226) < 1e-6
    assert abs(float(jax_prng.normal(42, 1, partitionable=False)[0]) + 0.18471177) < 1e-6


# ------------------------------------------------------------- 1. the reference's trajectories
def _fit_noise_factor(rhs, data, times, truth, dtype, n_grid=161, **kw):
    """Best ||oracle(f2) - reference|| / ||reference - closed form|| per series over the factor
    that follows loop iteration 1 (the step [0.1, 1.1])."""
    M = data.shape[0]
    grid = np.linspace(2.0, 10.0, n_grid)
    G = grid.size
    fo = np.full((M * G, 2), np.nan)
    fo[:, 1] = np.tile(grid, M)
    cfg = {**dict(rtol=1e-5, atol=1e-5, dt0=0.1), **kw}
    r = do.solve(rhs, np.repeat(data[:, :1], G, axis=0).astype(dtype), THETA_TRUE.astype(dtype),
                 np.zeros((0,), dtype), times.astype(dtype), dtype=dtype, factor_override=fo, **cfg)
    assert (r.status == 0).all()
    ys = r.ys[:, :, 0].astype(np.float64).reshape(M, G, -1)
    nref = np.linalg.norm(data - truth, axis=1)
    return (np.linalg.norm(ys - data[:, None, :], axis=2) / nref[:, None]).min(axis=1)


def test_oracle_reproduces_the_discretisation_error_of_the_reference_trajectories():
    data, times = _golden()
    assert data.shape == (120, 20)
    truth = closed_form(times[None, :], data[:, :1], THETA_TRUE[1], THETA_TRUE[0])
    eref = data - truth
    # the fingerprint is there: errors far above f32 rounding (ulp(300) = 3e-5)
    assert np.abs(eref).max() > 2e-2 and np.median(np.linalg.norm(eref, axis=1)) > 3e-4
    rhs, *_ = sy.make_rhs(*MENTEN[:4], MENTEN[4])

    ratio = _fit_noise_factor(rhs, data, times, truth, np.float32)
    # the oracle explains ~90 % of the reference's error vector in nearly every series
    # (floor: f32 rounding of both results ~ 0.05-0.1 of the typical error norm)
    assert np.median(ratio) < 0.15, np.median(ratio)
    assert (ratio < 0.25).mean() > 0.85, (ratio < 0.25).mean()
    assert ratio.max() < 0.7, ratio.max()

    # discriminating power: the same fit with another solver, or the tolerance off by 2x,
    # leaves most of the error unexplained
    # (measured: correct config median 0.105, 96 % of the series below 0.25;
    #  Dopri5 0.51 / 0 %, rtol x2 0.52 / 30 %, rtol / 2 0.39 / 23 %)
    for kw, floor in ((dict(solver="dopri5"), 0.4), (dict(rtol=2e-5, atol=2e-5), 0.4),
                      (dict(rtol=5e-6, atol=5e-6), 0.3)):
        wrong = _fit_noise_factor(rhs, data[::2], times, truth[::2], np.float32, n_grid=81, **kw)
        assert np.median(wrong) > floor, (kw, np.median(wrong))
        assert (wrong < 0.25).mean() < 0.5, (kw, (wrong < 0.25).mean())


def test_unfitted_f32_oracle_matches_the_reference_pattern_in_distribution():
    """Without any fitted factor the f32 oracle still lands on the reference's error vector for
    a large fraction of the series (where the rounding-noise factor happens to fall alike); the
    same statistic is asserted for the CUDA kernel in tests/test_gpu_reference_pins.py.
    Other configurations: <= 0.16 / <= 0.65 (measured: Dopri5 0.02 / 0.43, rtol 2e-5 0.15 / 0.65,
    rtol 5e-6 0.16 / 0.61, dt0 0.05 0.10 / 0.41, f64 arithmetic 0.08 / 0.01)."""
    data, times = _golden()
    truth = closed_form(times[None, :], data[:, :1], THETA_TRUE[1], THETA_TRUE[0])
    rhs, *_ = sy.make_rhs(*MENTEN[:4], MENTEN[4])
    f32 = np.float32
    r = do.solve(rhs, data[:, :1].astype(f32), THETA_TRUE.astype(f32), np.zeros((0,), f32),
                 times.astype(f32), rtol=1e-5, atol=1e-5, dt0=0.1, dtype=f32)
    cos = fingerprint_cosine(r.ys[:, :, 0].astype(np.float64), data, truth)
    assert (cos > 0.9).mean() > 0.3, (cos > 0.9).mean()
    assert np.median(cos) > 0.75, np.median(cos)
    # the C / OpenMP port that bench.py times as the CPU baseline (its own rounding: 40 %, 0.83)
    orc = COracle(*MENTEN[:4], MENTEN[4], dtype=f32)
    rc = orc.solve(data[:, :1], np.tile(THETA_TRUE, (data.shape[0], 1)), np.zeros((data.shape[0], 0)),
                   times, rtol=1e-5, atol=1e-5, dt0=0.1)
    cos_c = fingerprint_cosine(rc.ys[:, :, 0].astype(np.float64), data, truth)
    assert (cos_c > 0.9).mean() > 0.3 and np.median(cos_c) > 0.75, ((cos_c > 0.9).mean(), np.median(cos_c))


def fingerprint_cosine(ys, data, truth):
    e, eref = ys - truth, data - truth
    return (e * eref).sum(1) / (np.linalg.norm(e, axis=1) * np.linalg.norm(eref, axis=1) + 1e-300)


# ------------------------------------------------- 2. the reference's published NUTS posterior
KM_GRID = np.linspace(97.0, 109.0, 41)
VM_GRID = np.linspace(6.88, 7.30, 41)
SIG_GRID = np.linspace(1.9, 2.45, 56)


def posterior_moments(logp, theta, sig):
    """Moments of exp(logp [G, n_sigma]) on the tensor grid theta [G, 2] x sig."""
    w = np.exp(logp - logp.max())
    w /= w.sum()
    wt, ws = w.sum(1), w.sum(0)
    out = {}
    for name, x, wx in (("K_m", theta[:, 0], wt), ("v_max", theta[:, 1], wt), ("sigma", sig, ws)):
        m = (wx * x).sum()
        out[name] = (m, np.sqrt((wx * (x - m) ** 2).sum()))
    n = KM_GRID.size
    edge = max(wt.reshape(n, -1).sum(1)[[0, -1]].max(), wt.reshape(n, -1).sum(0)[[0, -1]].max(),
               ws[[0, -1]].max())
    return out, edge


def check_against_hmc_notebook(moments):
    for name, (mean, sd, mcse_mean, mcse_sd) in HMC_SUMMARY.items():
        got_mean, got_sd = moments[name]
        # 3 Monte-Carlo standard errors + the 3-decimal rounding of the printed table
        assert abs(got_mean - mean) < 3 * mcse_mean + 5e-4, (name, got_mean, mean)
        assert abs(got_sd - sd) < 3 * mcse_sd + 5e-4, (name, got_sd, sd)


def _theta_grid():
    KM, VM = np.meshgrid(KM_GRID, VM_GRID, indexing="ij")
    return np.stack([KM.ravel(), VM.ravel()], 1)


def _oracle_logp(sim, observed, likelihood, sig=SIG_GRID):
    """log joint on the grid: oracle likelihood terms + HalfNormal(yerrs=2) on sigma (uniform
    priors on K_m / v_max are constant inside their support)."""
    logp = np.empty((sim.shape[0], sig.size))
    for j, s in enumerate(sig):
        lp, _, _ = lo.logpdf_terms(observed[None, :], sim, s, likelihood)
        logp[:, j] = lp.sum(1) - 0.5 * (s / 2.0) ** 2
    return logp


def test_oracle_log_density_reproduces_the_posterior_of_the_hmc_notebook():
    data, times = _golden()
    M, T = data.shape
    theta = _theta_grid()
    G = theta.shape[0]
    # HMC.run -> SimulationConfig defaults rtol=atol=1e-5, dt0=0.1 (mcmc.py:73-79); f64 oracle:
    # the solver error (1e-5 relative) is far below sigma ~ 2
    orc = COracle(*MENTEN[:4], MENTEN[4], dtype=np.float64)
    r = orc.solve(np.tile(data[:, :1], (G, 1)), np.repeat(theta, M, axis=0), np.zeros((G * M, 0)),
                  times, rtol=1e-5, atol=1e-5, dt0=0.1)
    assert (r.status == 0).all()
    sim = r.ys[:, :, 0].reshape(G, M * T)

    observed = jax_prng.augment_multiplicative(data, 5e-2, seed=0).astype(np.float64).ravel()
    moments, edge = posterior_moments(_oracle_logp(sim, observed, "softlaplace"), theta, SIG_GRID)
    assert edge < 1e-3, edge                      # the grid holds the whole posterior
    check_against_hmc_notebook(moments)

    # discriminating power: the other likelihood, or the legacy PRNG layout (another noise
    # vector), are many standard errors away
    for lik, part in (("normal", True), ("softlaplace", False)):
        obs = jax_prng.augment_multiplicative(data, 5e-2, seed=0, partitionable=part)
        sig = np.linspace(1.9, 6.0, 56) if lik == "normal" else SIG_GRID
        wrong, _ = posterior_moments(_oracle_logp(sim, obs.astype(np.float64).ravel(), lik, sig),
                                     theta, sig)
        with pytest.raises(AssertionError):
            check_against_hmc_notebook(wrong)


# --------------------- 3. a second published posterior: ragged, NaN-masked, three observables
# examples/InhomogeneousData.ipynb cell 5 (4 chains x 1000 draws): mean, sd, mcse_mean, mcse_sd.
# Columns the table prints as 0.000 are bounded by the 3-decimal rounding (5e-4).
INACT_SUMMARY = {"K_M": (187.057, 5.462, 0.088, 0.062),
                 "k_cat": (1.098, 0.013, 5e-4, 5e-4),
                 "k_inact": (0.004, 0.000, 5e-4, 5e-4),
                 "sigma": (1.005, 0.082, 0.001, 0.001)}
INACT_GRIDS = (np.linspace(162.0, 220.0, 59), np.linspace(1.045, 1.155, 31),
               np.linspace(0.0030, 0.0054, 31), np.linspace(0.72, 1.42, 41))
INACT_GOLD = Path(__file__).parent / "golden" / "enzymeml_inactivation.npz"


def inactivation_model_and_dataset():
    """``ctx.from_enzymeml`` + cell 2 + ``dataset.pad()`` of the notebook, from the arrays of the
    EnzymeML document: three ODEs (Michaelis-Menten with first-order enzyme inactivation), three
    measurements of 19 points, the last one cut to 9 and NaN-padded back to 19."""
    import catalax_b200 as ctx
    from catalax_b200.dataset.measurement import Measurement
    from catalax_b200.mcmc import priors

    g = np.load(INACT_GOLD)
    states = [str(s) for s in g["states"]]
    model = ctx.Model(name="inactivation")
    model.add_state(**{s: s for s in states})
    for s, eq in zip(states, g["odes"]):
        model.add_ode(s, str(eq))
    model.parameters.k_cat.prior = priors.Uniform(low=0.1, high=2)
    model.parameters.K_M.prior = priors.Uniform(low=100, high=500)
    model.parameters.k_inact.prior = priors.Uniform(low=0.0001, high=0.01)
    assert model.get_parameter_order() == ["K_M", "k_cat", "k_inact"]
    dataset = ctx.Dataset.from_model(model)
    for i in range(g["data"].shape[0]):
        n = 19 if i < 2 else 9                          # species.data[:-10] of the last measurement
        dataset.add_measurement(Measurement(
            initial_conditions={s: float(g["y0"][i, j]) for j, s in enumerate(states)},
            data={s: g["data"][i, :n, j] for j, s in enumerate(states)}, time=g["times"][i, :n]))
    padded = dataset.pad()
    assert [len(m.time) for m in padded.measurements] == [19, 19, 19]
    return model, padded, states


def inactivation_moments(logp):
    """Moments of exp(logp) on the tensor grid INACT_GRIDS (C order)."""
    shape = tuple(g.size for g in INACT_GRIDS)
    w = np.exp(logp.reshape(shape) - logp.max())
    w /= w.sum()
    out, edge = {}, 0.0
    for ax, name in enumerate(("K_M", "k_cat", "k_inact", "sigma")):
        wx = w.sum(tuple(a for a in range(4) if a != ax))
        m = (wx * INACT_GRIDS[ax]).sum()
        out[name] = (m, np.sqrt((wx * (INACT_GRIDS[ax] - m) ** 2).sum()))
        edge = max(edge, wx[0], wx[-1])
    return out, edge


def check_against_inactivation_notebook(moments):
    for name, (mean, sd, mcse_mean, mcse_sd) in INACT_SUMMARY.items():
        got_mean, got_sd = moments[name]
        assert abs(got_mean - mean) < 3 * mcse_mean + 5e-4, (name, got_mean, mean)
        assert abs(got_sd - sd) < 3 * mcse_sd + 5e-4, (name, got_sd, sd)


def test_masked_three_state_posterior_of_the_inhomogeneous_data_notebook():
    """Host side of the log-density entry point (Dataset.pad, NaN mask, priors, the shared scalar
    noise site of the notebook's Catalax version) through ``BayesianModel.log_density``; the CUDA
    likelihood block is replaced by the oracle here (no GPU) and is the real kernel in
    tests/test_gpu_reference_pins.py."""
    import torch

    import catalax_b200 as ctx
    from catalax_b200.mcmc import BayesianModel, error

    ctx.enable_x64(True)
    try:
        model, padded, states = inactivation_model_and_dataset()
        bm = BayesianModel(model, padded, yerrs=2.0, error_model=error.HalfNormal(shared=True),
                           shard="none")
        data, times, y0s, consts = bm._arrays
        assert bm.mask.sum() == 3 * (19 + 19 + 9) and times[2, -1] == 90.0   # pad: 81, 82, ... 90
        odes = {s: str(model.odes[s].equation) for s in states}
        orc = COracle(states, ["K_M", "k_cat", "k_inact"], [], odes, dtype=np.float64)

        def oracle_block(theta, sigma):
            th, sg = theta.numpy(), sigma.numpy()
            G = th.shape[0]
            r = orc.solve(np.tile(y0s, (G, 1)), np.repeat(th, 3, axis=0), np.zeros((3 * G, 0)),
                          np.tile(times, (G, 1)), rtol=1e-5, atol=1e-5, dt0=0.1)
            assert (r.status == 0).all()
            ok = bm.mask.ravel()
            sim, obs = r.ys.reshape(G, -1)[:, ok], data.ravel()[ok]
            out = np.zeros((G, 1 + 3 + 3))
            lp, _, _ = lo.logpdf_terms(obs[None, :], sim, sg[:, :1], "softlaplace")
            out[:, 0] = lp.sum(1)
            return torch.as_tensor(out)

        bm._local_loglik = oracle_block
        mesh = np.meshgrid(*INACT_GRIDS[:3], indexing="ij")
        theta = np.stack([m.ravel() for m in mesh], 1)
        logp = np.empty((theta.shape[0], INACT_GRIDS[3].size))
        for j, s in enumerate(INACT_GRIDS[3]):
            lp, _, g_sg = bm.log_density(torch.as_tensor(theta), torch.full((theta.shape[0], 1), s,
                                                                            dtype=torch.float64))
            assert g_sg.shape == (theta.shape[0], 1)
            logp[:, j] = lp.numpy()
        moments, edge = inactivation_moments(logp)
        assert edge < 1e-3, edge
        check_against_inactivation_notebook(moments)
        with pytest.raises(ValueError):                 # the shared site is one scalar
            bm.log_density(torch.as_tensor(theta[:2]), torch.ones((2, 3), dtype=torch.float64))
    finally:
        ctx.enable_x64(False)
