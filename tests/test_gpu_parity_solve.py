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
              atol=1e-6, dt0=0.1, max_steps=4096, kernel=0, warp_form="auto"):
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle import diffrax_oracle as do, symbolic as sy

    states, params, consts_n, odes, reactions = model
    rhs, N, D, _ = sy.make_rhs(states, params, consts_n, odes, reactions)
    ref = do.solve(rhs, y0, theta, consts, ts, solver=solver, rtol=rtol, atol=atol, dt0=dt0,
                   max_steps=max_steps, dtype=dtype)
    field = cg.generate_field(cases.field_spec(model), warp_form=warp_form)
    m = engine.CompiledModel(field, np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=dtype), device=dev)
    out = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), solver=solver, in_axes=in_axes,
                  rtol=rtol, atol=atol, dt0=dt0, max_steps=max_steps, kernel=kernel)
    torch.cuda.synchronize()
    return ref, out


KERNELS = [1, 2]  # 1 = static thread-per-trajectory, 2 = persistent refill + cooperative saves


def _assert_close_f64(ys, ref_ys, rel=1e-6, atol=1e-6, abs_frac=1e-3):
    """x64 bound: |diff| <= 1e-6*|ref| + 1e-3*atol.  The absolute term only matters for
    components that have decayed below the solver's own atol, where both results are
    stability-limited noise of size ~atol (observed: 1.4e-10 on values of +-1e-6)."""
    np.testing.assert_array_equal(np.isinf(ys), np.isinf(ref_ys))
    fin = np.isfinite(ref_ys)
    bound = rel * np.abs(ref_ys[fin]) + abs_frac * atol
    assert (np.abs(ys[fin] - ref_ys[fin]) <= bound).all(), \
        float((np.abs(ys[fin] - ref_ys[fin]) / bound).max())


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("solver", ["tsit5", "dopri5"])
def test_mm_config1_f64(solver, kernel):
    """BASELINE config 1: README Michaelis-Menten, 2 rows, t1=100, nsteps=1000."""
    y0 = np.array([[10.0, 0.0, 100.0], [10.0, 0.0, 200.0]])
    theta = np.array([0.05, 0.1])
    ts = np.linspace(0, 100, 1000)
    ref, out = _run_pair(cases.michaelis_menten(), y0, theta, np.zeros((2, 0)), ts,
                         (0, None, 0, None), np.float64, solver=solver, rtol=1e-5, atol=1e-5,
                         kernel=kernel)
    ys = out.ys.cpu().numpy()
    assert (out.status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    np.testing.assert_array_equal(out.n_rejected.cpu().numpy(), ref.n_rejected)
    _assert_close_f64(ys, ref.ys, atol=1e-5)
    assert np.abs(ys - ref.ys).max() <= 1e-9


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("T", [16, 200])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_mm_batch_per_row_theta(dtype, T, kernel):
    """BASELINE config 2 distributions at a size the oracle finishes in seconds."""
    B = 2048 + 37  # ragged: not a multiple of the warp / block size
    y0, theta, consts = cases.mm_inputs(B, seed=0, dtype=dtype)
    ts = np.linspace(0, 100, T).astype(dtype)
    ref, out = _run_pair(cases.michaelis_menten(), y0, theta, consts, ts, (0, 0, 0, None), dtype,
                         kernel=kernel)
    ys = out.ys.cpu().numpy()
    assert (out.status.cpu().numpy() == ref.status).all() and (ref.status == 0).all()
    if dtype == np.float64:
        same = (out.n_accepted.cpu().numpy() == ref.n_accepted) & \
               (out.n_rejected.cpu().numpy() == ref.n_rejected)
        assert same.mean() > 0.999
        _assert_close_f64(ys, ref.ys)
    else:
        # f32 bound (stated): at rtol=atol=1e-6 the f32 error estimate is partly rounding noise,
        # so two implementations of the same algorithm take different accept/reject paths.
        # Measure both against an x64 truth: the CUDA result must be (a) within
        # 200*(atol + rtol*|y|) of it and (b) as accurate as the f32 oracle, statistically.
        from oracle import diffrax_oracle as do, symbolic as sy

        m = cases.michaelis_menten()
        rhs, *_ = sy.make_rhs(*m[:4], m[4])
        truth = do.solve(rhs, y0.astype(np.float64), theta.astype(np.float64),
                         consts.astype(np.float64), ts.astype(np.float64), rtol=1e-10,
                         atol=1e-10, dtype=np.float64, max_steps=100000).ys
        scale = 1e-6 + 1e-6 * np.abs(truth)
        e_gpu = np.abs(ys.astype(np.float64) - truth) / scale
        e_ref = np.abs(ref.ys.astype(np.float64) - truth) / scale
        assert e_gpu.max() <= 200.0, e_gpu.max()
        assert e_gpu.mean() <= 1.25 * e_ref.mean() + 0.05
        assert np.percentile(e_gpu, 99.9) <= 1.5 * np.percentile(e_ref, 99.9) + 1.0
        steps_gpu = (out.n_accepted + out.n_rejected).cpu().numpy().sum()
        steps_ref = (ref.n_accepted + ref.n_rejected).sum()
        assert abs(steps_gpu - steps_ref) / steps_ref < 0.02


@pytest.mark.parametrize("kernel,T,dtype", [(5, 384, np.float64), (0, 768, np.float64),
                                            (5, 1000, np.float32), (0, 320, np.float32)])
def test_warp_specialised_kernel_against_the_c_oracle(kernel, T, dtype):
    """The headline kernel (ctx_solve_v2: producer warps step, consumer warps save) compared with
    the oracle DIRECTLY -- explicitly (kernel 5) and as what ``kernel = 0`` selects for long shared
    grids (f32 >= 320 points, f64 >= 768, >= 65 536 rows) -- on BASELINE config 2's distributions
    at a size that engages every part of it (slot ring, guided queue, 8-point groups in f32).
    x64: 1e-6 relative at every save point and identical accepted / rejected counts; f32: the
    stated statistical bound against an x64 truth, next to the f32 oracle."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle.c_oracle import COracle

    B = 65536 + 1237      # ragged tail
    model = cases.michaelis_menten()
    y0, theta, consts = cases.mm_inputs(B, seed=11, dtype=dtype)
    ts = np.linspace(0, 100, T).astype(dtype)
    orc = COracle(*model[:4], model[4], dtype=dtype)
    ref = orc.solve(y0, theta, consts, ts, rtol=1e-6, atol=1e-6, dt0=0.1, max_steps=4096)
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    out = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), in_axes=(0, 0, 0, None), rtol=1e-6,
                  atol=1e-6, dt0=0.1, max_steps=4096, kernel=kernel)
    torch.cuda.synchronize()
    assert engine.last_solve_kernel() == "ctx_solve_v2"
    ys = out.ys.cpu().numpy()
    assert (out.status.cpu().numpy() == 0).all() and (ref.status == 0).all()
    if dtype == np.float64:
        np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
        np.testing.assert_array_equal(out.n_rejected.cpu().numpy(), ref.n_rejected)
        # |diff| <= 1e-6 |y| + 0.1 atol at EVERY save point of all 66 773 rows.  The absolute term
        # matters only for rows in the stability-limited regime after substrate depletion (thousands
        # of steps on a component that is +-atol noise): there the rounding differences between FMA
        # (GPU) and non-contracted (oracle) arithmetic are amplified to a few 1e-8 absolute (measured:
        # 4.6e-8, i.e. 5e-6 relative on values ~1e-2).  Everywhere else the agreement is ~1e-12: the
        # 99.9th percentile of the relative difference is asserted at 1e-9.
        d = np.abs(ys - ref.ys)
        assert (d <= 1e-6 * np.abs(ref.ys) + 0.1 * 1e-6).all(), float((d / (1e-6 * np.abs(ref.ys) + 1e-7)).max())
        big = np.abs(ref.ys) > 1.0
        q = float(np.quantile((d[big] / np.abs(ref.ys[big]))[::7], 0.999))
        print(f"v2 x64: max abs diff {float(d.max()):.2e}, 99.9th percentile rel diff {q:.2e}")
        assert q <= 1e-9, q
    else:
        sub = slice(0, 8192)
        truth = COracle(*model[:4], model[4], dtype=np.float64).solve(
            y0[sub].astype(np.float64), theta[sub].astype(np.float64), consts[sub].astype(np.float64),
            ts.astype(np.float64), rtol=1e-10, atol=1e-10, dt0=0.1, max_steps=100000).ys
        scale = 1e-6 + 1e-6 * np.abs(truth)
        e_gpu = np.abs(ys[sub].astype(np.float64) - truth) / scale
        e_ref = np.abs(ref.ys[sub].astype(np.float64) - truth) / scale
        assert e_gpu.max() <= 200.0, e_gpu.max()
        assert e_gpu.mean() <= 1.25 * e_ref.mean() + 0.05
        assert np.percentile(e_gpu, 99.9) <= 1.5 * np.percentile(e_ref, 99.9) + 1.0
        steps_gpu = int((out.n_accepted + out.n_rejected).sum().item())
        steps_ref = int((ref.n_accepted.astype(np.int64) + ref.n_rejected).sum())
        assert abs(steps_gpu - steps_ref) / steps_ref < 0.02


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("solver,dt0", [("tsit5", 0.1), ("dopri5", 0.1), ("euler", 0.01), ("heun", 0.02)])
def test_nonautonomous_all_solvers_f64(solver, dt0, kernel):
    rng = np.random.default_rng(5)
    B = 64
    y0 = rng.uniform(0.2, 2.0, (B, 4))
    theta = rng.uniform(0.1, 1.0, (3,))
    consts = rng.uniform(0.0, 0.5, (B, 1))
    ts = np.sort(rng.uniform(0, 6.0, (B, 12)), axis=1)
    ts[:, 0] = 0.0
    ref, out = _run_pair(cases.nonautonomous(), y0, theta, consts, ts, (0, None, 0, 0),
                         np.float64, solver=solver, dt0=dt0, rtol=1e-7, atol=1e-9,
                         max_steps=100000, kernel=kernel)
    ys = out.ys.cpu().numpy()
    assert (out.status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    _assert_close_f64(ys, ref.ys, atol=1e-9 if solver in ("tsit5", "dopri5") else 1e-6)


@pytest.mark.parametrize("kernel", KERNELS)
def test_stiff_rows_fail_like_the_reference(kernel):
    """SURVEY's original K_m range: some rows exhaust max_steps; status + inf pattern must match."""
    B = 1024
    y0, theta, consts = cases.mm_inputs(B, seed=0, dtype=np.float64, km_range=(0.01, 1.0))
    ts = np.linspace(0, 100, 16)
    ref, out = _run_pair(cases.michaelis_menten(), y0, theta, consts, ts, (0, 0, 0, None),
                         np.float64, kernel=kernel)
    assert (ref.status == 1).sum() > 10
    np.testing.assert_array_equal(out.status.cpu().numpy(), ref.status)
    # thousands of stability-limited steps: the depleted component is +-atol noise that
    # amplifies rounding differences chaotically -> compare it to the full atol
    _assert_close_f64(out.ys.cpu().numpy(), ref.ys, abs_frac=1.0)


@pytest.mark.parametrize("kernel", KERNELS)
def test_degenerate_t0_equals_t1(kernel):
    y0 = np.array([[10.0, 0.0, 100.0], [10.0, 1.0, 50.0]])
    ts = np.zeros((2, 3))
    ref, out = _run_pair(cases.michaelis_menten(), y0, np.array([0.05, 0.1]), np.zeros((2, 0)),
                         ts, (0, None, 0, 0), np.float64, kernel=kernel)
    ys = out.ys.cpu().numpy()
    np.testing.assert_array_equal(ys, np.broadcast_to(y0[:, None, :], ys.shape))
    np.testing.assert_array_equal(ys, ref.ys)


@pytest.mark.parametrize("kernel", KERNELS)
def test_max_steps_failure_fills_inf(kernel):
    """throw=False semantics (simulation.py:255-257): unsaved slots are inf, status says why."""
    y0 = np.array([[10.0, 0.0, 100.0]])
    theta = np.array([0.05, 0.1])
    ts = np.linspace(0, 100, 50)
    ref, out = _run_pair(cases.michaelis_menten(), y0, theta, np.zeros((1, 0)), ts,
                         (0, None, 0, None), np.float64, rtol=1e-8, atol=1e-8, max_steps=20,
                         kernel=kernel)
    ys = out.ys.cpu().numpy()
    assert out.status.cpu().numpy()[0] == 1 == ref.status[0]
    np.testing.assert_array_equal(np.isinf(ys), np.isinf(ref.ys))
    assert np.isinf(ys).any() and np.isfinite(ys[0, 0]).all()
    fin = np.isfinite(ref.ys)
    np.testing.assert_allclose(ys[fin], ref.ys[fin], rtol=1e-9)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("solver", ["tsit5", "dopri5", "heun"])
def test_sensitivities_match_oracle_f64(solver, kernel):
    """ctx_solve_sens = vmap(jax.jacobian(solve)) (simulation.py:279-323): discrete tangents
    w.r.t. parameters and initial values, layout [B,T,S,D]."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle import diffrax_oracle as do, symbolic as sy

    model = cases.nonautonomous()
    states, params, consts_n, odes, reactions = model
    rng = np.random.default_rng(8)
    B, S, P = 96, 4, 3
    y0 = rng.uniform(0.3, 1.5, (B, S)); th = rng.uniform(0.2, 0.9, (B, P))
    c = rng.uniform(0.0, 0.4, (B, 1)); ts = np.linspace(0, 5, 11)
    dt0 = 0.1 if solver != "heun" else 0.02
    kw = dict(solver=solver, rtol=1e-6, atol=1e-8, dt0=dt0, max_steps=100000)
    rhs, N, D, init_aug = sy.make_rhs(states, params, consts_n, odes, reactions, sens_theta=True,
                                      sens_y0=True)
    ref = do.solve(rhs, y0, th, c, ts, dtype=np.float64, n_err=S, init_aug=init_aug(B, np.float64), **kw)
    field = cg.generate_field(cases.field_spec(model), sens_theta=True, sens_y0=True)
    m = engine.CompiledModel(field, "float64")
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    out = m.solve(tt(y0), tt(th), tt(c), tt(ts), in_axes=(0, 0, 0, None), tangents=True,
                  kernel=kernel, **kw)
    torch.cuda.synchronize()
    ys, sens = out.ys.cpu().numpy(), out.sens.cpu().numpy()
    assert sens.shape == (B, len(ts), S, P + S)
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    _assert_close_f64(ys, ref.ys[..., :S], atol=1e-8)
    ref_sens = ref.ys[..., S:].reshape(B, len(ts), D, S).transpose(0, 1, 3, 2)
    scale = np.abs(ref_sens) + 1e-6 * np.abs(ref_sens).max()
    assert (np.abs(sens - ref_sens) <= 1e-6 * scale).all()
    # initial values: d y(0) / d y0 = identity, d y(0) / d theta = 0
    np.testing.assert_array_equal(sens[:, 0, :, P:], np.broadcast_to(np.eye(S), (B, S, S)))
    np.testing.assert_array_equal(sens[:, 0, :, :P], 0.0)


@pytest.mark.parametrize("kernel", KERNELS + [4, 0])
@pytest.mark.parametrize("per_row_k", [True, False])
def test_mass_action_network_config3(per_row_k, kernel):
    """BASELINE configs[2] at test size: 32 states / 64 reactions (reaction form, stoichiometric
    sums), per-row or shared rate constants, T=32 on [0,10], rtol=atol=1e-6."""
    rng = np.random.default_rng(1)
    B = 96
    y0 = rng.uniform(0.1, 2.0, (B, 32))
    k = np.exp(rng.uniform(np.log(0.01), np.log(1.0), (B, 64) if per_row_k else (64,)))
    ts = np.linspace(0, 10, 32)
    ref, out = _run_pair(cases.mass_action_network(32), y0, k, np.zeros((B, 0)), ts,
                         (0, 0 if per_row_k else None, 0, None), np.float64, kernel=kernel)
    ys = out.ys.cpu().numpy()
    assert (out.status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    _assert_close_f64(ys, ref.ys)
    # conservation: first-order steps conserve total mass, the A+B->C couplings remove one unit
    assert (ys.sum(axis=2)[:, -1] <= ys.sum(axis=2)[:, 0] + 1e-9).all()


# ----------------------------------------------------------------------------- K1w (kernel 4)
@pytest.mark.parametrize("solver,dt0", [("tsit5", 0.1), ("dopri5", 0.1), ("euler", 0.01), ("heun", 0.02)])
def test_warp_kernel_general_field_all_solvers_f64(solver, dt0):
    """ctx_solve_w with a field that has 4 structural classes, t, constants and *_init symbols,
    per-row time grids (forced lane-parallel form; 4 of 32 lanes own a state)."""
    rng = np.random.default_rng(5)
    B = 67
    y0 = rng.uniform(0.2, 2.0, (B, 4))
    theta = rng.uniform(0.1, 1.0, (3,))
    consts = rng.uniform(0.0, 0.5, (B, 1))
    ts = np.sort(rng.uniform(0, 6.0, (B, 12)), axis=1)
    ts[:, 0] = 0.0
    ref, out = _run_pair(cases.nonautonomous(), y0, theta, consts, ts, (0, None, 0, 0),
                         np.float64, solver=solver, dt0=dt0, rtol=1e-7, atol=1e-9,
                         max_steps=100000, kernel=4, warp_form=True)
    ys = out.ys.cpu().numpy()
    assert (out.status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    np.testing.assert_array_equal(out.n_rejected.cpu().numpy(), ref.n_rejected)
    _assert_close_f64(ys, ref.ys, atol=1e-9 if solver in ("tsit5", "dopri5") else 1e-6)


@pytest.mark.parametrize("n_states", [40, 64])
def test_warp_kernel_more_states_than_lanes(n_states):
    """S > 32: every lane owns two states (the second slot ragged for S=40); T=70 > 32 so the
    requested times of one step span more than one prefetch round."""
    rng = np.random.default_rng(11)
    B = 33
    y0 = rng.uniform(0.1, 2.0, (B, n_states))
    k = np.exp(rng.uniform(np.log(0.01), np.log(1.0), (B, 2 * n_states)))
    ts = np.linspace(0, 10, 70)
    ref, out = _run_pair(cases.mass_action_network(n_states), y0, k, np.zeros((B, 0)), ts,
                         (0, 0, 0, None), np.float64, kernel=4)
    assert (out.status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    _assert_close_f64(out.ys.cpu().numpy(), ref.ys)


def test_warp_kernel_failure_semantics_and_degenerate_grid():
    """max_steps exhaustion -> status 1 + inf fill; t0 == t1 -> y0 everywhere (kernel 4)."""
    model = cases.mass_action_network(32)
    rng = np.random.default_rng(3)
    y0 = rng.uniform(0.1, 2.0, (5, 32))
    k = np.exp(rng.uniform(np.log(0.01), np.log(1.0), (64,)))
    ts = np.linspace(0, 10, 32)
    ref, out = _run_pair(model, y0, k, np.zeros((5, 0)), ts, (0, None, 0, None), np.float64,
                         rtol=1e-9, atol=1e-9, max_steps=12, kernel=4)
    ys = out.ys.cpu().numpy()
    np.testing.assert_array_equal(out.status.cpu().numpy(), ref.status)
    assert (ref.status == 1).all()
    np.testing.assert_array_equal(np.isinf(ys), np.isinf(ref.ys))
    fin = np.isfinite(ref.ys)
    np.testing.assert_allclose(ys[fin], ref.ys[fin], rtol=1e-9)
    ref, out = _run_pair(model, y0, k, np.zeros((5, 0)), np.zeros((5, 3)), (0, None, 0, 0),
                         np.float64, kernel=4)
    np.testing.assert_array_equal(out.ys.cpu().numpy(),
                                  np.broadcast_to(y0[:, None, :], (5, 3, 32)))


def test_warp_kernel_f32_config3_statistical_bound():
    """f32 (stated bound, as for kernels 1/2): error against an x64 truth within
    200*(atol+rtol*|y|) and statistically no worse than the f32 oracle."""
    from oracle import diffrax_oracle as do, symbolic as sy

    model = cases.mass_action_network(32)
    rng = np.random.default_rng(1)
    B = 256
    y0 = rng.uniform(0.1, 2.0, (B, 32)).astype(np.float32)
    k = np.exp(rng.uniform(np.log(0.01), np.log(1.0), (B, 64))).astype(np.float32)
    ts = np.linspace(0, 10, 32).astype(np.float32)
    ref, out = _run_pair(model, y0, k, np.zeros((B, 0), np.float32), ts, (0, 0, 0, None),
                         np.float32, kernel=4)
    rhs, *_ = sy.make_rhs(*model[:4], model[4])
    truth = do.solve(rhs, y0.astype(np.float64), k.astype(np.float64), np.zeros((B, 0)),
                     ts.astype(np.float64), rtol=1e-10, atol=1e-10, dtype=np.float64,
                     max_steps=100000).ys
    ys = out.ys.cpu().numpy()
    scale = 1e-6 + 1e-6 * np.abs(truth)
    e_gpu = np.abs(ys.astype(np.float64) - truth) / scale
    e_ref = np.abs(ref.ys.astype(np.float64) - truth) / scale
    assert (out.status.cpu().numpy() == 0).all()
    assert e_gpu.max() <= 200.0, e_gpu.max()
    assert e_gpu.mean() <= 1.25 * e_ref.mean() + 0.05
    assert np.percentile(e_gpu, 99.9) <= 1.5 * np.percentile(e_ref, 99.9) + 1.0
