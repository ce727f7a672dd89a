"""CUDA path vs the oracle on the same seeded inputs (bit-level algorithm parity).

Tolerances (stated, per north_star): x64: <= 1e-6 relative at every save point (observed
~1e-12); f32: |diff| <= 50*(atol + rtol*|y|) + 2e-5*|y| (an accept/reject flip caused by a
1-ulp difference changes a trajectory at the level of the solver tolerance).
"""
import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _run_pair(model, y0, theta, consts, ts, in_axes, dtype, solver="tsit5", rtol=1e-6,
              atol=1e-6, dt0=0.1, max_steps=4096, kernel=0):
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle import diffrax_oracle as do, symbolic as sy

    states, params, consts_n, odes, reactions = model
    rhs, N, D, _ = sy.make_rhs(states, params, consts_n, odes, reactions)
    ref = do.solve(rhs, y0, theta, consts, ts, solver=solver, rtol=rtol, atol=atol, dt0=dt0,
                   max_steps=max_steps, dtype=dtype)
    field = cg.generate_field(cases.field_spec(model))
    m = engine.CompiledModel(field, np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=dtype), device=dev)
    out = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), solver=solver, in_axes=in_axes,
                  rtol=rtol, atol=atol, dt0=dt0, max_steps=max_steps, kernel=kernel)
    torch.cuda.synchronize()
    return ref, out


@pytest.mark.parametrize("solver", ["tsit5", "dopri5"])
def test_mm_config1_f64(solver):
    """BASELINE config 1: README Michaelis-Menten, 2 rows, t1=100, nsteps=1000."""
    y0 = np.array([[10.0, 0.0, 100.0], [10.0, 0.0, 200.0]])
    theta = np.array([0.05, 0.1])
    ts = np.linspace(0, 100, 1000)
    ref, out = _run_pair(cases.michaelis_menten(), y0, theta, np.zeros((2, 0)), ts,
                         (0, None, 0, None), np.float64, solver=solver, rtol=1e-5, atol=1e-5)
    ys = out.ys.cpu().numpy()
    assert (out.status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    np.testing.assert_array_equal(out.n_rejected.cpu().numpy(), ref.n_rejected)
    rel = np.abs(ys - ref.ys) / np.maximum(np.abs(ref.ys), 1e-300)
    assert rel[np.abs(ref.ys) > 1e-12].max() <= 1e-6
    assert np.abs(ys - ref.ys).max() <= 1e-9


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_mm_batch_per_row_theta(dtype):
    """BASELINE config 2 distributions at a size the oracle finishes in seconds."""
    B = 2048
    y0, theta, consts = cases.mm_inputs(B, seed=0, dtype=dtype)
    ts = np.linspace(0, 100, 16).astype(dtype)
    ref, out = _run_pair(cases.michaelis_menten(), y0, theta, consts, ts, (0, 0, 0, None), dtype)
    ys = out.ys.cpu().numpy()
    assert (out.status.cpu().numpy() == ref.status).all()
    if dtype == np.float64:
        same = (out.n_accepted.cpu().numpy() == ref.n_accepted) & \
               (out.n_rejected.cpu().numpy() == ref.n_rejected)
        assert same.mean() > 0.999
        denom = np.maximum(np.abs(ref.ys), 1e-6)
        assert (np.abs(ys - ref.ys) / denom).max() <= 1e-6
    else:
        bound = 50 * (1e-6 + 1e-6 * np.abs(ref.ys)) + 2e-5 * np.abs(ref.ys)
        assert (np.abs(ys - ref.ys) <= bound).all()
        steps_gpu = (out.n_accepted + out.n_rejected).cpu().numpy().sum()
        steps_ref = (ref.n_accepted + ref.n_rejected).sum()
        assert abs(steps_gpu - steps_ref) / steps_ref < 0.02


@pytest.mark.parametrize("solver,dt0", [("tsit5", 0.1), ("dopri5", 0.1), ("euler", 0.01), ("heun", 0.02)])
def test_nonautonomous_all_solvers_f64(solver, dt0):
    rng = np.random.default_rng(5)
    B = 64
    y0 = rng.uniform(0.2, 2.0, (B, 4))
    theta = rng.uniform(0.1, 1.0, (3,))
    consts = rng.uniform(0.0, 0.5, (B, 1))
    ts = np.sort(rng.uniform(0, 6.0, (B, 12)), axis=1)
    ts[:, 0] = 0.0
    ref, out = _run_pair(cases.nonautonomous(), y0, theta, consts, ts, (0, None, 0, 0),
                         np.float64, solver=solver, dt0=dt0, rtol=1e-7, atol=1e-9,
                         max_steps=100000)
    ys = out.ys.cpu().numpy()
    assert (out.status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    assert (np.abs(ys - ref.ys) / np.maximum(np.abs(ref.ys), 1e-3)).max() <= 1e-6


def test_max_steps_failure_fills_inf():
    """throw=False semantics (simulation.py:255-257): unsaved slots are inf, status says why."""
    y0 = np.array([[10.0, 0.0, 100.0]])
    theta = np.array([0.05, 0.1])
    ts = np.linspace(0, 100, 50)
    ref, out = _run_pair(cases.michaelis_menten(), y0, theta, np.zeros((1, 0)), ts,
                         (0, None, 0, None), np.float64, rtol=1e-8, atol=1e-8, max_steps=20)
    ys = out.ys.cpu().numpy()
    assert out.status.cpu().numpy()[0] == 1 == ref.status[0]
    np.testing.assert_array_equal(np.isinf(ys), np.isinf(ref.ys))
    assert np.isinf(ys).any() and np.isfinite(ys[0, 0]).all()
    fin = np.isfinite(ref.ys)
    np.testing.assert_allclose(ys[fin], ref.ys[fin], rtol=1e-9)
