"""MCMC data preparation against the reference's REAL ``catalax.mcmc.mcmc``
(tests/golden/make_mcmc_prep_golden.py): ``MCMCConfig.to_simulation_config`` incl. the
``t1 = max_steps`` quirk, ``_extract_dataset_components`` on a padded ragged data set with a
non-observable state and a missing observation, the finite mask."""
import json
from pathlib import Path

import numpy as np

G = np.load(Path(__file__).parent / "golden" / "mcmc_prep_reference.npz")


def _facade():
    import warnings

    import catalax_b200 as ctx

    m = ctx.Model(name="Michaelis-Menten")
    m.add_state(S="Substrate", E="Enzyme", P="Product")
    m.add_constant(I="Inhibitor")
    m.add_assignment("v", "k_cat * E * S / (K_m * (1 + I / K_i) + S)")
    m.add_ode("S", "-v")
    m.add_ode("P", "v")
    m.add_ode("E", "0", observable=False)
    ds = ctx.Dataset.from_model(m)
    rng = np.random.default_rng(3)
    for i, (s0, e0, inh, T) in enumerate(((100.0, 10.0, 0.0, 6), (200.0, 5.0, 1.5, 6), (50.0, 20.0, 3.0, 4))):
        t = np.linspace(0, 50, T)
        data = {"S": s0 * np.exp(-t / 40) + rng.normal(0, 0.1, T), "P": s0 * (1 - np.exp(-t / 40))}
        if i == 1:
            data["S"][2] = np.nan
        ds.add_measurement(ctx.Measurement(id=f"m{i}", initial_conditions={"S": s0, "E": e0, "P": 0.0, "I": inh},
                                           time=t, data=data))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return m, ds.pad()


def test_mcmc_config_to_simulation_config_quirk():
    from catalax_b200.mcmc import MCMCConfig
    from catalax_b200.solvers import solver_name

    for key, cfg in (("sim_config", MCMCConfig(num_warmup=10, num_samples=10, dt0=0.05, max_steps=5000)),
                     ("default_sim_config", MCMCConfig(num_warmup=1, num_samples=1))):
        want = json.loads(str(G[key]))
        got = cfg.to_simulation_config()
        assert got.t1 == want["t1"] and got.t0 == want["t0"] and got.dt0 == want["dt0"]     # t1 = max_steps
        assert got.rtol == want["rtol"] and got.atol == want["atol"]
        assert got.max_steps == want["max_steps"] and got.throw == want["throw"]          # max_steps stays 4096
        assert solver_name(got.solver) == want["solver"]


def test_extract_dataset_components_and_mask():
    import catalax_b200 as ctx
    from catalax_b200.mcmc import extract_dataset_components

    ctx.enable_x64(True)
    try:
        m, ds = _facade()
        assert m.get_observable_state_order() == list(G["observable_order"])
        assert m.get_state_order() == list(G["state_order"])
        data, times, y0s, constants = extract_dataset_components(ds, m)
        np.testing.assert_allclose(np.asarray(data), G["data"], rtol=1e-15, equal_nan=True)
        np.testing.assert_allclose(np.asarray(times), G["times"], rtol=1e-15, equal_nan=True)
        np.testing.assert_allclose(np.asarray(y0s), G["y0s"], rtol=1e-15)          # ALL states, E included
        np.testing.assert_allclose(np.asarray(constants), G["constants"], rtol=1e-15)
        np.testing.assert_array_equal(np.isfinite(np.asarray(data)), G["mask"])
        shapes = json.loads(str(G["shapes"]))
        assert list(np.shape(y0s)) == shapes["y0s"] and list(np.shape(data)) == shapes["data"]
    finally:
        ctx.enable_x64(False)
