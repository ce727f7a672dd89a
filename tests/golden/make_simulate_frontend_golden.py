"""Golden records of the reference's REAL ``Model.simulate`` front end, run
in the build container.

The reference's ``Model.simulate`` -> ``Simulation._prepare_func`` -> ``jax.vmap(_simulate_system)``
-> ``diffrax.diffeqsolve`` call chain runs UNMODIFIED (catalax/model/model.py:511-745,
catalax/tools/simulation.py:160-355) through the import hook of make_model_facade_golden.py, with
``jax.vmap`` -> a loop over the mapped axes, ``eqx.filter_jit`` -> identity and ``diffrax`` -> a
recording adapter: every ``diffeqsolve`` call is logged (the operands that reach the solver:
t0, t1, dt0, y0, args, SaveAt.ts, controller tolerances, max_steps, throw, solver) and answered by
the parity oracle's solve of the reference's own vector field (the stack object it was given).
Recorded: the per-lane solver operands and the arrays / Dataset that ``simulate`` returns.
Output: tests/golden/simulate_frontend_reference.npz
"""
import json
import sys
import types
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from oracle import diffrax_oracle as do   # noqa: E402

OUT = Path(__file__).parent / "simulate_frontend_reference.npz"
CALLS = []


# ------------------------------------------------------------------ diffrax stand-in
class _Solver:
    def __init__(self):
        pass


class Tsit5(_Solver):
    name = "tsit5"


class Dopri5(_Solver):
    name = "dopri5"


class Euler(_Solver):
    name = "euler"


class Heun(_Solver):
    name = "heun"


class ODETerm:
    def __init__(self, vector_field):
        self.vector_field = vector_field


class SaveAt:
    def __init__(self, ts=None, **kw):
        self.ts = ts


class PIDController:
    def __init__(self, rtol, atol, **kw):
        self.rtol, self.atol = rtol, atol


class ConstantStepSize:
    rtol = atol = None


def diffeqsolve(terms, solver, t0, t1, dt0, y0, args, saveat, stepsize_controller, max_steps, throw, **kw):
    parameters, constants, inits = args
    ts = np.asarray(saveat.ts, dtype=np.float64)
    CALLS.append(dict(solver=solver.name, t0=float(t0), t1=float(t1), dt0=float(dt0),
                      y0=np.array(y0, dtype=float), parameters=np.array(parameters, dtype=float),
                      constants=np.array(constants, dtype=float), inits=np.array(inits, dtype=float), ts=ts.copy(),
                      rtol=stepsize_controller.rtol, atol=stepsize_controller.atol, max_steps=int(max_steps),
                      throw=bool(throw)))
    vf = terms.vector_field

    def rhs(t, Y, th, c, yinit):      # oracle convention: batched over the leading axis
        return np.stack([np.asarray(vf(t[b], Y[b], (th[b], c[b], yinit[b])), dtype=float) for b in range(Y.shape[0])])

    r = do.solve(rhs, np.asarray(y0, dtype=float)[None], np.asarray(parameters, dtype=float)[None],
                 np.asarray(constants, dtype=float)[None], ts[None], solver=solver.name,
                 rtol=stepsize_controller.rtol or 1e-5, atol=stepsize_controller.atol or 1e-5, dt0=float(dt0),
                 max_steps=int(max_steps), dtype=np.float64)
    if throw and r.status[0] != 0:
        raise RuntimeError("max_steps reached")
    return types.SimpleNamespace(ys=r.ys[0])


dfx = types.ModuleType("diffrax")          # in sys.modules BEFORE the reference is imported
for obj in (Tsit5, Dopri5, Euler, Heun, ODETerm, SaveAt, PIDController, ConstantStepSize):
    setattr(dfx, obj.__name__, obj)
dfx.diffeqsolve = diffeqsolve
dfx.AbstractSolver = _Solver
dfx.__getattr__ = lambda name: type(name, (), {})      # anything else the reference only names
sys.modules["diffrax"] = dfx

import make_model_facade_golden as mk   # noqa: E402  (stubs + the reference Model)


def vmap(f, in_axes=0, **kw):
    def run(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [f(*[(np.take(a, i, axis=ax) if ax is not None else a) for a, ax in zip(args, axes)]) for i in range(n)]
        return np.stack(outs)
    return run


jax = sys.modules["jax"]
jax.vmap = vmap
jax.jit = lambda f=None, **k: (f if f is not None else (lambda g: g))
eqx = sys.modules["equinox"]
eqx.filter_jit = lambda f=None, **k: (f if f is not None else (lambda g: g))

Model = mk.Model
SimulationConfig = sys.modules["catalax.model.simconfig"].SimulationConfig
InAxes = sys.modules["catalax.model.inaxes"].InAxes
Dataset = sys.modules["catalax.dataset.dataset"].Dataset


def mm_model():
    m = Model(name="Michaelis-Menten")
    m.add_state(S="Substrate", E="Enzyme", P="Product")
    m.add_constant(I="Inhibitor")
    m.add_assignment("v", "k_cat * E * S / (K_m * (1 + I / K_i) + S)")
    m.add_ode("S", "-v")
    m.add_ode("P", "v")
    m.add_ode("E", "0")
    m.parameters.K_m.value, m.parameters.k_cat.value, m.parameters.K_i.value = 30.0, 0.4, 2.0
    return m


def main():
    out = {}
    m = mm_model()
    ds = Dataset.from_model(m)
    ds.add_initial(S=100.0, E=10.0, P=0.0, I=0.0)
    ds.add_initial(S=200.0, E=5.0, P=1.0, I=1.5)
    ds.add_initial(S=50.0, E=20.0, P=0.0, I=3.0)

    def run(tag, **kw):
        CALLS.clear()
        res = m.simulate(ds, **kw)
        out[f"{tag}/n_calls"] = np.array(len(CALLS))
        for k in ("y0", "parameters", "constants", "inits", "ts"):
            out[f"{tag}/{k}"] = np.stack([c[k] for c in CALLS])
        out[f"{tag}/scalars"] = np.array(json.dumps([{k: c[k] for k in ("solver", "t0", "t1", "dt0", "rtol", "atol",
                                                                      "max_steps", "throw")} for c in CALLS]))
        return res

    # 1. config.nsteps: shared grid, in_axes = InAxes.Y0
    res = run("nsteps", config=SimulationConfig(t0=0, t1=60, nsteps=13, dt0=0.1, rtol=1e-6, atol=1e-6))
    out["nsteps/data"] = np.stack([np.stack([meas.data[s] for s in m.get_state_order()], axis=-1)
                                   for meas in res.measurements])
    out["nsteps/time"] = np.stack([np.asarray(meas.time) for meas in res.measurements])
    out["nsteps/initial_conditions"] = np.array(json.dumps([dict(meas.initial_conditions) for meas in res.measurements]))
    # 2. explicit 2-D saveat (per-series grids) + return_array + Dopri5
    saveat = np.stack([np.linspace(0, 40, 9), np.linspace(0, 80, 9), np.array([0, 1, 2, 4, 8, 16, 32, 64, 90.0])])
    t_arr, ys = run("saveat2d", config=SimulationConfig(t1=90, solver=Dopri5, dt0=0.05), saveat=saveat, return_array=True)
    out["saveat2d/ys"], out["saveat2d/t"] = np.asarray(ys), np.asarray(t_arr)
    # 3. explicit parameters argument
    t_arr, ys = run("parameters", config=SimulationConfig(t1=30, nsteps=5), parameters=np.array([1.0, 60.0, 0.9]),
                    return_array=True)
    out["parameters/ys"] = np.asarray(ys)
    # 4. Model.predict (model.py:569-603): config from the data set's own time spans
    #    (Dataset.to_config: per-measurement t0 / t1 arrays -> a 2-D linspace grid), and use_times
    Measurement = sys.modules["catalax.dataset.measurement"].Measurement
    pds = Dataset.from_model(m)
    for i, (s0, e0, inh, t) in enumerate(((100.0, 10.0, 0.0, np.linspace(0, 40, 6)),
                                          (200.0, 5.0, 1.5, np.linspace(5, 90, 6)))):
        pds.add_measurement(Measurement(id=f"p{i}", initial_conditions={"S": s0, "E": e0, "P": 0.0, "I": inh}, time=t,
                                        data={"S": np.zeros(6), "P": np.zeros(6), "E": np.zeros(6)}))
    out["predict/times_in"] = np.stack([np.asarray(meas.time) for meas in pds.measurements])
    for tag, kw in (("predict", dict(n_steps=7)), ("predict_use_times", dict(use_times=True))):
        CALLS.clear()
        res = m.predict(pds, **kw)
        out[f"{tag}/n_calls"] = np.array(len(CALLS))
        for k in ("y0", "parameters", "constants", "inits", "ts"):
            out[f"{tag}/{k}"] = np.stack([c[k] for c in CALLS])
        out[f"{tag}/scalars"] = np.array(json.dumps([{k: c[k] for k in ("solver", "t0", "t1", "dt0", "rtol", "atol",
                                                                       "max_steps", "throw")} for c in CALLS]))
        out[f"{tag}/data"] = np.stack([np.stack([meas.data[s] for s in m.get_state_order()], axis=-1)
                                       for meas in res.measurements])
        out[f"{tag}/time"] = np.stack([np.asarray(meas.time) for meas in res.measurements])
    out["model"] = np.array(json.dumps({"state_order": m.get_state_order(), "parameter_order": m.get_parameter_order(),
                                        "constants_order": m.get_constants_order()}))
    np.savez_compressed(OUT, **out)
    for k in sorted(out):
        print(k, out[k].shape if out[k].ndim else str(out[k])[:150])


if __name__ == "__main__":
    main()
