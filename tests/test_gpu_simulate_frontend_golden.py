"""``Model.simulate`` end to end against the reference's REAL front end
(tests/golden/make_simulate_frontend_golden.py runs the reference's ``Model.simulate`` ->
``Simulation._prepare_func`` -> ``vmap(_simulate_system)`` -> ``diffeqsolve`` chain unmodified over a
recording diffrax adapter backed by the parity oracle): the operands that reach the solver -- per
lane y0 / parameters / constants / inits / save times, solver, tolerances, dt0, max_steps -- and the
arrays / Dataset that come back."""
import json
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

G = np.load(Path(__file__).parent / "golden" / "simulate_frontend_reference.npz")


def _model_and_dataset():
    import catalax_b200 as ctx

    m = ctx.Model(name="Michaelis-Menten")
    m.add_state(S="Substrate", E="Enzyme", P="Product")
    m.add_constant(I="Inhibitor")
    m.add_assignment("v", "k_cat * E * S / (K_m * (1 + I / K_i) + S)")
    m.add_ode("S", "-v")
    m.add_ode("P", "v")
    m.add_ode("E", "0")
    m.parameters.K_m.value, m.parameters.k_cat.value, m.parameters.K_i.value = 30.0, 0.4, 2.0
    ds = ctx.Dataset.from_model(m)
    ds.add_initial(S=100.0, E=10.0, P=0.0, I=0.0)
    ds.add_initial(S=200.0, E=5.0, P=1.0, I=1.5)
    ds.add_initial(S=50.0, E=20.0, P=0.0, I=3.0)
    return m, ds


@pytest.fixture
def recorded(monkeypatch):
    from catalax_b200 import engine

    calls = []
    real = engine.CompiledModel.solve

    def spy(self, y0, theta, consts, ts, **kw):
        calls.append(dict(y0=y0.cpu().numpy(), theta=theta.cpu().numpy(), consts=consts.cpu().numpy(),
                          ts=ts.cpu().numpy(), kw={k: v for k, v in kw.items() if k != "order"}))
        return real(self, y0, theta, consts, ts, **kw)

    monkeypatch.setattr(engine.CompiledModel, "solve", spy)
    return calls


def _check_operands(call, tag):
    """The batched operands of ctx_solve, lane by lane, against what diffeqsolve received."""
    want = {k: G[f"{tag}/{k}"] for k in ("y0", "parameters", "constants", "inits", "ts")}
    scal = json.loads(str(G[f"{tag}/scalars"]))
    B = int(G[f"{tag}/n_calls"])
    axes = call["kw"]["in_axes"]
    lane = lambda a, ax, b: a[b] if ax == 0 else a
    for b in range(B):
        np.testing.assert_allclose(lane(call["y0"], axes[0], b), want["y0"][b], rtol=1e-15)
        np.testing.assert_allclose(lane(call["y0"], axes[0], b), want["inits"][b], rtol=1e-15)   # args[2] = y0
        np.testing.assert_allclose(lane(call["theta"], axes[1], b), want["parameters"][b], rtol=1e-15)
        np.testing.assert_allclose(lane(call["consts"], axes[2], b), want["constants"][b], rtol=1e-15)
        np.testing.assert_allclose(lane(call["ts"], axes[3], b), want["ts"][b], rtol=1e-15)
        s = scal[b]
        assert call["kw"]["solver"] == s["solver"] and s["t0"] == 0.0          # simulation.py:262: t0 = 0
        assert call["kw"]["rtol"] == s["rtol"] and call["kw"]["atol"] == s["atol"]
        assert call["kw"]["dt0"] == s["dt0"] and call["kw"]["max_steps"] == s["max_steps"]
        assert float(lane(call["ts"], axes[3], b)[-1]) == s["t1"]              # t1 = time[-1]


def test_simulate_matches_the_reference_front_end(recorded):
    import catalax_b200 as ctx
    from catalax_b200.solvers import Dopri5

    ctx.enable_x64(True)
    try:
        m, ds = _model_and_dataset()
        order = json.loads(str(G["model"]))
        assert m.get_state_order() == order["state_order"] and m.get_parameter_order() == order["parameter_order"]
        assert m.get_constants_order() == order["constants_order"]

        # 1. config.nsteps -> a shared grid, in_axes = InAxes.Y0, a Dataset comes back
        res = m.simulate(ds, ctx.SimulationConfig(t0=0, t1=60, nsteps=13, dt0=0.1, rtol=1e-6, atol=1e-6))
        assert recorded[-1]["kw"]["in_axes"] == (0, None, 0, None)
        _check_operands(recorded[-1], "nsteps")
        data = np.stack([np.stack([meas.data[s] for s in m.get_state_order()], axis=-1) for meas in res.measurements])
        np.testing.assert_allclose(data, G["nsteps/data"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(np.stack([meas.time for meas in res.measurements]), G["nsteps/time"], rtol=1e-15)
        assert [dict(meas.initial_conditions) for meas in res.measurements] == json.loads(str(G["nsteps/initial_conditions"]))

        # 2. explicit 2-D saveat (per-series grids), Dopri5, arrays back
        saveat = np.stack([np.linspace(0, 40, 9), np.linspace(0, 80, 9), np.array([0, 1, 2, 4, 8, 16, 32, 64, 90.0])])
        t_arr, ys = m.simulate(ds, ctx.SimulationConfig(t1=90, solver=Dopri5, dt0=0.05), saveat=saveat,
                               return_array=True)
        assert recorded[-1]["kw"]["in_axes"] == (0, None, 0, 0)
        _check_operands(recorded[-1], "saveat2d")
        np.testing.assert_allclose(np.asarray(t_arr), G["saveat2d/t"], rtol=1e-15)
        np.testing.assert_allclose(np.asarray(ys), G["saveat2d/ys"], rtol=1e-9, atol=1e-11)

        # 3. the `parameters=` argument replaces the model's values
        _, ys = m.simulate(ds, ctx.SimulationConfig(t1=30, nsteps=5), parameters=np.array([1.0, 60.0, 0.9]),
                           return_array=True)
        _check_operands(recorded[-1], "parameters")
        np.testing.assert_allclose(np.asarray(ys), G["parameters/ys"], rtol=1e-9, atol=1e-11)

        # 4. Model.predict: the grid comes from the data set's own time spans (per-measurement t0 / t1
        #    -> a 2-D linspace), or the measured times themselves (use_times); the solve still starts
        #    at t = 0 (simulation.py:262) even where a series' first time is 5
        pds = ctx.Dataset.from_model(m)
        for i, (s0, e0, inh, t) in enumerate(((100.0, 10.0, 0.0, np.linspace(0, 40, 6)),
                                              (200.0, 5.0, 1.5, np.linspace(5, 90, 6)))):
            pds.add_measurement(ctx.Measurement(id=f"p{i}", initial_conditions={"S": s0, "E": e0, "P": 0.0, "I": inh},
                                                time=t, data={"S": np.zeros(6), "P": np.zeros(6), "E": np.zeros(6)}))
        for tag, kw in (("predict", dict(n_steps=7)), ("predict_use_times", dict(use_times=True))):
            res = m.predict(pds, **kw)
            assert recorded[-1]["kw"]["in_axes"] == (0, None, 0, 0)
            _check_operands(recorded[-1], tag)
            data = np.stack([np.stack([meas.data[s] for s in m.get_state_order()], axis=-1) for meas in res.measurements])
            np.testing.assert_allclose(data, G[f"{tag}/data"], rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(np.stack([meas.time for meas in res.measurements]), G[f"{tag}/time"], rtol=1e-14)
    finally:
        ctx.enable_x64(False)
