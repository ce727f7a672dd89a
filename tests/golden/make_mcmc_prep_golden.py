"""Golden records of the reference's REAL MCMC data preparation, run in the
build container: ``catalax.mcmc.mcmc`` imported unmodified (import hook + diffrax adapter of
make_simulate_frontend_golden.py; numpyro is a stub module, it is only named here):
``MCMCConfig(...).to_simulation_config()`` (the ``t1 = max_steps`` quirk, defaults),
``_prepare_mcmc_data`` -> ``_extract_dataset_components`` (observable data, times, FULL initial
states incl. non-observable ones, constants), the finite mask, the shapes record, and the
``in_axes = (0, None, 0, 0)`` solve function it builds (evaluated once through the recording
adapter).  Output: tests/golden/mcmc_prep_reference.npz
"""
import importlib
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_simulate_frontend_golden as fe   # noqa: E402  (stubs, diffrax adapter, reference Model)

OUT = Path(__file__).parent / "mcmc_prep_reference.npz"


def main():
    mcmc = importlib.import_module("catalax.mcmc.mcmc")
    Measurement = sys.modules["catalax.dataset.measurement"].Measurement
    m = fe.mm_model()
    m.odes["E"].observable = False                      # a non-observable state
    ds = fe.Dataset.from_model(m)
    rng = np.random.default_rng(3)
    for i, (s0, e0, inh, T) in enumerate(((100.0, 10.0, 0.0, 6), (200.0, 5.0, 1.5, 6), (50.0, 20.0, 3.0, 4))):
        t = np.linspace(0, 50, T)
        data = {"S": s0 * np.exp(-t / 40) + rng.normal(0, 0.1, T), "P": s0 * (1 - np.exp(-t / 40))}
        if i == 1:
            data["S"][2] = np.nan                        # a missing observation
        ds.add_measurement(Measurement(id=f"m{i}", initial_conditions={"S": s0, "E": e0, "P": 0.0, "I": inh},
                                       time=t, data=data))
    ds = ds.pad()                                        # ragged -> NaN padded (InhomogeneousData.ipynb)
    cfg = mcmc.MCMCConfig(num_warmup=10, num_samples=10, dt0=0.05, max_steps=5000)
    sim_cfg = cfg.to_simulation_config()
    default_cfg = mcmc.MCMCConfig(num_warmup=1, num_samples=1).to_simulation_config()
    prep = mcmc._prepare_mcmc_data(ds, m, surrogate=None, config=sim_cfg)
    out = {"data": np.asarray(prep.data, dtype=float), "times": np.asarray(prep.times, dtype=float),
           "y0s": np.asarray(prep.y0s, dtype=float), "constants": np.asarray(prep.constants, dtype=float),
           "mask": np.asarray(prep.mask, dtype=bool),
           "observable_order": np.array(m.get_observable_state_order()),
           "state_order": np.array(m.get_state_order()),
           "sim_config": np.array(json.dumps({k: (getattr(sim_cfg, k) if k != "solver" else sim_cfg.solver.name)
                                              for k in ("t0", "t1", "dt0", "rtol", "atol", "max_steps", "throw", "solver")})),
           "default_sim_config": np.array(json.dumps({k: (getattr(default_cfg, k) if k != "solver" else default_cfg.solver.name)
                                                      for k in ("t0", "t1", "dt0", "rtol", "atol", "max_steps", "throw", "solver")})),
           "shapes": np.array(json.dumps({k: list(getattr(prep.shapes, k)) for k in ("y0s", "constants", "times", "data")}))}
    # the solve function the model body calls: (y0s, theta, constants, times) with in_axes (0, None, 0, 0)
    fe.CALLS.clear()
    theta = np.array([2.0, 30.0, 0.4])
    times_safe = np.where(np.isnan(out["times"]), np.nanmax(out["times"]), out["times"])
    states = prep.sim_func(out["y0s"], theta, out["constants"], times_safe)
    out["sim_states"] = np.asarray(states, dtype=float)
    out["sim_theta"] = theta
    out["sim_times"] = times_safe
    out["sim_calls"] = np.array(json.dumps([{k: c[k] for k in ("solver", "t0", "t1", "dt0", "rtol", "atol", "max_steps", "throw")}
                                            for c in fe.CALLS]))
    np.savez_compressed(OUT, **out)
    for k in sorted(out):
        print(k, out[k].shape if out[k].ndim else str(out[k])[:160])


if __name__ == "__main__":
    main()
