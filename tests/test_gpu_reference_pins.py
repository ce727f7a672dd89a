"""The CUDA path against OUTPUTS OF THE REFERENCE ITSELF (see tests/test_reference_pins.py for
what the two artefacts are and how the oracle is pinned on them).  Everything goes through the
drop-in facade: ``Model.simulate`` (f32, default config) and ``BayesianModel.log_density``."""
from pathlib import Path

import numpy as np
import pytest

import catalax_b200 as ctx
from oracle import jax_prng
from tests.test_reference_pins import (HMC_SUMMARY, KM_GRID, SIG_GRID, THETA_TRUE, VM_GRID,
                                       check_against_hmc_notebook, closed_form,
                                       fingerprint_cosine, posterior_moments)

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLD = Path(__file__).parent / "golden" / "croissant_menten.npz"


def _menten_model():
    # examples/GenerateData.ipynb cell 3 / examples/HMC.ipynb cell 3
    model = ctx.Model(name="Simple menten model")
    model.add_state("s1")
    model.add_ode("s1", "- (v_max * s1) / ( K_m + s1)")
    model.parameters["v_max"].value = 7.0
    model.parameters["K_m"].value = 100.0
    return model


def test_simulate_f32_lands_on_the_reference_trajectories():
    """GenerateData.ipynb cells 4-5 through the facade on the GPU, f32, default config.  The
    reference's 120 outputs carry their discretisation error (100x f32 rounding) as a fingerprint
    of the accepted steps; the CUDA result must carry the same one wherever the rounding-noise
    factor after the step [0.1, 1.1] falls alike (numpy f32 oracle: 40 % of the series above
    0.9 cosine, median 0.86; any other solver / tolerance / dt0: <= 16 %, <= 0.65)."""
    g = np.load(GOLD)
    data, times = g["data"].astype(np.float64), g["times"]
    ctx.enable_x64(False)
    model = _menten_model()
    dataset = ctx.Dataset.from_model(model)
    for s0 in g["data"][:, 0]:
        dataset.add_initial(s1=float(s0))
    _, result = model.simulate(dataset=dataset, config=ctx.SimulationConfig(t0=0, t1=100),
                               saveat=times, return_array=True)
    ys = np.asarray(result.cpu() if hasattr(result, "cpu") else result, dtype=np.float64)[:, :, 0]
    assert ys.shape == data.shape and np.isfinite(ys).all()
    truth = closed_form(times.astype(np.float64)[None, :], data[:, :1], THETA_TRUE[1], THETA_TRUE[0])
    # same accuracy class as the reference's own run ...
    rms, rms_ref = np.sqrt(((ys - truth) ** 2).mean()), np.sqrt(((data - truth) ** 2).mean())
    assert 0.4 < rms / rms_ref < 2.5, (rms, rms_ref)
    assert np.abs(ys - data).max() < 6e-2          # 2 x the reference's own worst error
    # ... and the same error vectors
    cos = fingerprint_cosine(ys, data, truth)
    print(f"fingerprint: frac(cos>0.9)={(cos > 0.9).mean():.3f} median={np.median(cos):.3f} "
          f"rms/rms_ref={rms / rms_ref:.3f}")
    assert (cos > 0.9).mean() > 0.25, (cos > 0.9).mean()
    assert np.median(cos) > 0.7, np.median(cos)


@pytest.mark.parametrize("kernel", [1, 2])
def test_cuda_kernel_reproduces_the_reference_trajectories_up_to_one_factor(kernel, monkeypatch):
    """The fitted form of the pin above, for the CUDA kernels: the factor that follows the step
    [0.1, 1.1] is f32 rounding noise in ANY implementation (tests/test_reference_pins.py), so it
    is scanned per series through the kernels' test hook (CTX_DBG_FACTOR_*: factor from an extra,
    unused constant) exactly as ``factor_override`` does for the oracle.  The CUDA kernel must
    explain the reference's error vectors as well as the oracle does (oracle: median residual
    0.105, 96 % of the series below 0.25) -- and not with another solver."""
    import sympy as sp

    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    g = np.load(GOLD)
    data, times = g["data"].astype(np.float64), g["times"].astype(np.float64)
    M = data.shape[0]
    truth = closed_form(times[None, :], data[:, :1], THETA_TRUE[1], THETA_TRUE[0])
    nref = np.linalg.norm(data - truth, axis=1)
    monkeypatch.setenv("CTX_B200_NVRTC_FLAGS", "-DCTX_DBG_FACTOR_CONST=0 -DCTX_DBG_FACTOR_STEP=1")
    s1, Km, vmax = sp.symbols("s1 K_m v_max")
    spec = cg.FieldSpec(["s1"], ["K_m", "v_max"], ["second_factor"], {"s1": -(vmax * s1) / (Km + s1)}, [])
    m = engine.CompiledModel(cg.generate_field(spec), "float32")
    dev = torch.device("cuda:0")
    grid = np.linspace(2.0, 10.0, 161)
    G = grid.size
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)

    def fit(solver):
        out = m.solve(tt(np.repeat(data[:, :1], G, axis=0)), tt(THETA_TRUE), tt(np.tile(grid, M)[:, None]),
                      tt(times), solver=solver, in_axes=(0, None, 0, None), rtol=1e-5, atol=1e-5,
                      dt0=0.1, kernel=kernel)
        assert int((out.status != 0).sum()) == 0
        ys = out.ys[:, :, 0].cpu().numpy().astype(np.float64).reshape(M, G, -1)
        return (np.linalg.norm(ys - data[:, None, :], axis=2) / nref[:, None]).min(axis=1)

    ratio = fit("tsit5")
    print(f"kernel {kernel}: fitted residual median {np.median(ratio):.3f}, "
          f"frac < 0.25: {(ratio < 0.25).mean():.3f}, max {ratio.max():.3f}")
    assert np.median(ratio) < 0.15, np.median(ratio)
    assert (ratio < 0.25).mean() > 0.85, (ratio < 0.25).mean()
    assert ratio.max() < 0.7, ratio.max()
    wrong = fit("dopri5")
    assert np.median(wrong) > 0.4 and (wrong < 0.25).mean() < 0.5, (np.median(wrong), (wrong < 0.25).mean())


def test_log_density_reproduces_the_posterior_of_the_hmc_notebook():
    """HMC.ipynb cells 3-6: priors Uniform(1e-6, 200) / Uniform(1e-6, 1e3), yerrs=2.0, default
    SoftLaplace likelihood and HalfNormal(yerrs) noise prior; the log joint evaluated by the CUDA
    likelihood kernel on a (K_m, v_max, sigma) grid of 94 136 chains x 120 series and integrated
    must give the posterior mean / sd the reference's NUTS run printed, to 3 MCSE."""
    from catalax_b200.mcmc import BayesianModel, priors

    g = np.load(GOLD)
    M, T = g["data"].shape
    # cell 4: the f32 data set, augmented through the facade exactly as the notebook does
    ctx.enable_x64(False)
    raw = ctx.Dataset.from_jax_arrays(["s1"], g["data"][:, :, None], np.tile(g["times"], (M, 1)),
                                      g["data"][:, :1])
    ds = raw.augment(n_augmentations=1, seed=0, sigma=5e-2, append=False, multiplicative=True)
    observed = np.stack([m.data["s1"] for m in ds.measurements])
    assert observed.dtype == np.float32
    np.testing.assert_array_equal(observed, jax_prng.augment_multiplicative(g["data"], 5e-2, seed=0))
    ctx.enable_x64(True)
    try:
        model = _menten_model()
        model.parameters["v_max"].prior = priors.Uniform(low=1e-6, high=200.0)
        model.parameters["K_m"].prior = priors.Uniform(low=1e-6, high=1e3)
        assert model.get_parameter_order() == ["K_m", "v_max"]
        bm = BayesianModel(model, ds, yerrs=2.0)
        assert bm.likelihood == "softlaplace"
        KM, VM, SG = np.meshgrid(KM_GRID, VM_GRID, SIG_GRID, indexing="ij")
        dev = torch.device("cuda:0")
        theta = torch.as_tensor(np.stack([KM.ravel(), VM.ravel()], 1), device=dev)
        sigma = torch.as_tensor(SG.reshape(-1, 1), device=dev)
        logp, g_th, g_sg = bm.log_density(theta, sigma)
        torch.cuda.synchronize()
        logp = logp.cpu().numpy().reshape(KM_GRID.size * VM_GRID.size, SIG_GRID.size)
        assert np.isfinite(logp).all()
        th_grid = np.stack([KM[:, :, 0].ravel(), VM[:, :, 0].ravel()], 1)
        moments, edge = posterior_moments(logp, th_grid, SIG_GRID)
        print("posterior (CUDA log-density):", {k: (round(m, 4), round(s, 4)) for k, (m, s) in moments.items()},
              "reference:", HMC_SUMMARY)
        assert edge < 1e-3, edge
        check_against_hmc_notebook(moments)
        # the gradient is the derivative of the pinned density: central differences along the
        # K_m axis of the grid (spacing 0.3) at the grid's interior, relative to the gradient scale
        lp3 = logp.reshape(KM_GRID.size, VM_GRID.size, SIG_GRID.size)
        gk = g_th[:, 0].cpu().numpy().reshape(lp3.shape)
        fd = (lp3[2:] - lp3[:-2]) / (KM_GRID[2] - KM_GRID[0])
        assert np.abs(fd - gk[1:-1]).max() < 5e-3 * np.abs(gk).max()   # oracle: 3e-4
    finally:
        ctx.enable_x64(False)


def test_masked_three_state_log_density_reproduces_the_inhomogeneous_data_notebook():
    """InhomogeneousData.ipynb cells 2-5 (three-state inactivation model from the EnzymeML
    document, last measurement cut and NaN-padded, one noise scalar shared by the three
    observables): the CUDA likelihood kernel on 2.3 M (K_M, k_cat, k_inact, sigma) chains x 3
    ragged series, integrated, against the summary the reference's 4-chain NUTS run printed."""
    from catalax_b200.mcmc import BayesianModel, error
    from tests.test_reference_pins import (INACT_GRIDS, INACT_SUMMARY, check_against_inactivation_notebook,
                                           inactivation_model_and_dataset, inactivation_moments)

    ctx.enable_x64(True)
    try:
        model, padded, _ = inactivation_model_and_dataset()
        bm = BayesianModel(model, padded, yerrs=2.0, error_model=error.HalfNormal(shared=True))
        mesh = np.meshgrid(*INACT_GRIDS, indexing="ij")
        dev = torch.device("cuda:0")
        theta = torch.as_tensor(np.stack([m.ravel() for m in mesh[:3]], 1), device=dev)
        sigma = torch.as_tensor(mesh[3].reshape(-1, 1), device=dev)
        logp, g_th, g_sg = bm.log_density(theta, sigma)
        torch.cuda.synchronize()
        assert g_sg.shape == (theta.shape[0], 1) and torch.isfinite(logp).all()
        moments, edge = inactivation_moments(logp.cpu().numpy())
        print("posterior (CUDA log-density):", {k: (round(m, 5), round(s, 5)) for k, (m, s) in moments.items()},
              "reference:", INACT_SUMMARY)
        assert edge < 1e-3, edge
        check_against_inactivation_notebook(moments)
        # d logp / d sigma of the shared site = central differences along the sigma axis
        shape = tuple(g.size for g in INACT_GRIDS)
        lp4 = logp.cpu().numpy().reshape(shape)
        gs = g_sg[:, 0].cpu().numpy().reshape(shape)
        fd = (lp4[..., 2:] - lp4[..., :-2]) / (INACT_GRIDS[3][2] - INACT_GRIDS[3][0])
        assert np.abs(fd - gs[..., 1:-1]).max() < 1e-2 * np.abs(gs).max()
    finally:
        ctx.enable_x64(False)
