"""Edge cases of the solve seam against the oracle: every in_axes combination the reference's
InAxes enum produces (inaxes.py:5-27), single-point and repeated-time grids, negative / early
save times (Appendix A: integration always starts at t0 = 0), the unbatched call, ragged sizes."""
import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _pair(model, y0, theta, consts, ts, in_axes, kernel, **kw):
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle import diffrax_oracle as do, symbolic as sy

    rhs, *_ = sy.make_rhs(*model[:4], model[4])
    ref = do.solve(rhs, y0, theta, consts, ts, dtype=np.float64, **kw)
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), "float64")
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)
    out = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), in_axes=in_axes, kernel=kernel, **kw)
    torch.cuda.synchronize()
    return ref, out


def _close(ys, ref, atol=1e-5):
    np.testing.assert_array_equal(np.isinf(ys), np.isinf(ref))
    fin = np.isfinite(ref)
    assert (np.abs(ys[fin] - ref[fin]) <= 1e-6 * np.abs(ref[fin]) + 1e-3 * atol).all()


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("in_axes", [(0, None, 0, None), (None, 0, None, None), (None, None, 0, None),
                                     (None, None, None, 0), (0, 0, 0, 0), (0, None, 0, 0)])
def test_every_in_axes_combination(in_axes, kernel):
    rng = np.random.default_rng(12)
    B = 37
    model = cases.menten_fixture()            # 1 state, 2 parameters, 1 constant
    y0 = rng.uniform(50, 150, (B, 1)) if in_axes[0] == 0 else np.array([100.0])
    th = np.stack([rng.uniform(50, 150, B), rng.uniform(-2, -0.5, B)], 1) if in_axes[1] == 0 else np.array([100.0, -1.0])
    c = rng.uniform(5, 15, (B, 1)) if in_axes[2] == 0 else np.array([10.0])
    ts = np.sort(rng.uniform(0, 20, (B, 9)), axis=1) if in_axes[3] == 0 else np.linspace(0, 20, 9)
    ref, out = _pair(model, y0, th, c, ts, in_axes, kernel, rtol=1e-5, atol=1e-5)
    assert out.ys.shape == (B, 9, 1)
    np.testing.assert_array_equal(out.n_accepted.cpu().numpy(), ref.n_accepted)
    _close(out.ys.cpu().numpy(), ref.ys)


@pytest.mark.parametrize("kernel", [1, 2])
def test_single_time_point_and_repeated_times(kernel):
    model = cases.michaelis_menten()
    y0 = np.array([[10.0, 0.0, 100.0], [12.0, 1.0, 80.0]])
    th = np.array([40.0, 0.3])
    ref, out = _pair(model, y0, th, np.zeros((2, 0)), np.array([7.5]), (0, None, 0, None), kernel)
    _close(out.ys.cpu().numpy(), ref.ys)
    ts = np.array([0.0, 0.0, 2.0, 2.0, 2.0, 9.0, 9.0])
    ref, out = _pair(model, y0, th, np.zeros((2, 0)), ts, (0, None, 0, None), kernel)
    ys = out.ys.cpu().numpy()
    _close(ys, ref.ys)
    np.testing.assert_array_equal(ys[:, 2], ys[:, 3])
    np.testing.assert_array_equal(ys[:, 0], y0)


@pytest.mark.parametrize("kernel", [1, 2])
def test_save_times_before_t0_are_extrapolated_from_the_first_step(kernel):
    """config.t0 only shapes the grid (model.py:664-665); the solve starts at 0 (simulation.py:262),
    so earlier times are evaluated with the first step's interpolant."""
    model = cases.michaelis_menten()
    y0 = np.array([[10.0, 0.0, 100.0]])
    ts = np.linspace(-0.05, 5.0, 12)
    ref, out = _pair(model, y0, np.array([40.0, 0.3]), np.zeros((1, 0)), ts, (0, None, 0, None), kernel)
    _close(out.ys.cpu().numpy(), ref.ys)
    assert out.ys[0, 0, 2].item() > 100.0      # extrapolated backwards: more substrate than y0


def test_unbatched_call_and_ragged_sizes_through_the_facade():
    import catalax_b200 as ctx
    from catalax_b200.tools.simulation import Simulation

    ctx.enable_x64(True)
    try:
        model = ctx.Model(name="m")
        model.add_state("s1")
        model.add_constant("e")
        model.add_ode("s1", "kcat * e * s1 / (k_m + s1)")
        sim_input = model.sim_input
        cfg = ctx.SimulationConfig(t1=10, nsteps=5)
        f, _ = Simulation(sim_input=sim_input, config=cfg)._prepare_func(in_axes=None)
        ys = f(np.array([100.0]), np.array([100.0, -1.0]), np.array([10.0]), np.linspace(0, 10, 5))
        assert ys.shape == (5, 1) and ys[0, 0] == 100.0 and ys[-1, 0] < 100.0
        for B in (1, 31, 33, 129):
            f, _ = Simulation(sim_input=sim_input, config=cfg)._prepare_func(in_axes=(0, None, 0, None))
            out = f(np.full((B, 1), 100.0), np.array([100.0, -1.0]), np.full((B, 1), 10.0),
                    np.linspace(0, 10, 5))
            assert out.shape == (B, 5, 1)
            np.testing.assert_allclose(out, np.broadcast_to(ys, (B, 5, 1)), rtol=1e-12)
    finally:
        ctx.enable_x64(False)


def test_long_segments_with_unaligned_rows_f32_and_f64():
    """The warp-wide 128-bit store path needs 16-byte aligned groups of 4 points; T*S*w that is
    not a multiple of 16 exercises its scalar fallback, T=1000 the vector path."""
    model = cases.michaelis_menten()
    rng = np.random.default_rng(3)
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from oracle import c_oracle

    for dtype in (np.float32, np.float64):
        co = c_oracle.COracle(*model[:4], model[4], dtype=dtype)
        m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), np.dtype(dtype).name)
        for T in (1000, 333, 50):
            B = 70
            y0, th, c = cases.mm_inputs(B, seed=T, dtype=dtype)
            ts = np.linspace(0, 100, T).astype(dtype)
            ref = co.solve(y0, th, c, ts, rtol=1e-6, atol=1e-6)
            dev = torch.device("cuda:0")
            tt = lambda a: torch.as_tensor(a, device=dev)
            out = m.solve(tt(y0), tt(th), tt(c), tt(ts), in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, kernel=2)
            ys = out.ys.cpu().numpy().astype(np.float64)
            if dtype == np.float64:
                _close(ys, ref.ys, atol=1e-6)
            else:
                scale = 1e-6 + 1e-6 * np.abs(ref.ys)
                assert np.median(np.abs(ys - ref.ys) / scale) < 2.0
                assert (np.abs(ys - ref.ys) <= 500 * scale).all()


@pytest.mark.parametrize("T,shared_ts", [(640, True), (517, True), (600, False), (1001, True)])
def test_wide_save_variant_is_bit_identical(T, shared_ts, monkeypatch):
    """Long f32 output grids (T >= 512) take the kernel variant whose lanes evaluate 8 requested
    times per pass instead of 4 (ctx_api.cu launch_solve).  It is the same arithmetic in a
    different schedule: results must be bit-identical to the 4-point variant, incl. ragged T,
    per-row grids and failed rows (inf fill)."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    model = cases.michaelis_menten()
    B = 3001
    y0, theta, consts = cases.mm_inputs(B, seed=5, dtype=np.float32)
    theta[:7, 0] = 1e-3                      # stiff rows: exhaust max_steps -> inf fill
    ts = np.linspace(0, 100, T, dtype=np.float32)
    if not shared_ts:
        ts = np.tile(ts, (B, 1)) * np.linspace(0.9, 1.1, B, dtype=np.float32)[:, None]
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), "float32")
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    outs = []
    for wide in ("0", "1"):
        monkeypatch.setenv("CTX_B200_WIDE_SAVE", wide)
        o = m.solve(tt(y0), tt(theta), tt(consts), tt(ts),
                    in_axes=(0, 0, 0, None if shared_ts else 0), rtol=1e-6, atol=1e-6, max_steps=512)
        torch.cuda.synchronize()
        outs.append((o.ys.cpu().numpy(), o.status.cpu().numpy(), o.n_accepted.cpu().numpy()))
    assert (outs[0][1] != 0).any() and (outs[0][1] == 0).any()
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    np.testing.assert_array_equal(outs[0][2], outs[1][2])
    np.testing.assert_array_equal(outs[0][0], outs[1][0])


@pytest.mark.parametrize("S", [1, 2, 4, 5])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_store_paths_agree_bitwise_for_every_state_count(S, dtype, monkeypatch):
    """The save path picks its store instructions from S and the dtype (256-bit whole-sector
    stores, parity-shifted for odd S in f32; 8-point groups for long f32 grids).  Whatever the
    path, the VALUES are the same arithmetic: compare with a build restricted to 4-point groups
    and 128-bit stores, on a short ragged grid and on a long one, with one constant state."""
    import sympy as sp

    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    states = [f"x{i}" for i in range(S)]
    params = [f"k{i}" for i in range(S)]
    odes = {}
    for i in range(S):
        e = -sp.Symbol(params[i]) * sp.Symbol(states[i])
        if i > 0:
            e = e + sp.Float(0.5) * sp.Symbol(states[i - 1])
        odes[states[i]] = e
    if S >= 4:
        odes[states[S - 1]] = sp.Integer(0)          # a constant state (CTX_CONST_MASK)
    model = (states, params, [], odes, [])
    rng = np.random.default_rng(S)
    B = 777
    y0 = rng.uniform(0.5, 2.0, (B, S)).astype(dtype)
    theta = rng.uniform(0.05, 1.0, (B, S)).astype(dtype)
    consts = np.zeros((B, 0), dtype)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    for T in (37, 1001):
        ts = np.linspace(0, 20, T).astype(dtype)
        res = []
        for flags, wide in (("", None), ("-DCTX_ST256=0 -DCTX_GP=4", "0")):
            monkeypatch.setenv("CTX_B200_NVRTC_FLAGS", flags)
            if wide is None:
                monkeypatch.delenv("CTX_B200_WIDE_SAVE", raising=False)
            else:
                monkeypatch.setenv("CTX_B200_WIDE_SAVE", wide)
            m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), np.dtype(dtype).name)
            o = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), in_axes=(0, 0, 0, None), rtol=1e-6,
                        atol=1e-6, kernel=2)
            torch.cuda.synchronize()
            assert (o.status == 0).all()
            res.append(o.ys.cpu().numpy())
        np.testing.assert_array_equal(res[0], res[1])
        # ... and they are the solution: x0(t) = x0(0) exp(-k0 t)
        want = y0[:, None, 0] * np.exp(-theta[:, None, 0].astype(np.float64) * ts[None, :].astype(np.float64))
        tol = 2e-4 if dtype == np.float32 else 2e-5
        assert np.abs(res[0][:, :, 0] - want).max() <= tol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("S,solver,B,T", [(3, "tsit5", 3001, 37), (3, "tsit5", 70000, 1001),
                                          (1, "dopri5", 517, 800), (2, "tsit5", 40, 5),
                                          (4, "dopri5", 9000, 300)])
def test_warp_specialised_kernel_is_bit_identical_to_the_persistent_kernel(S, solver, B, T, dtype):
    """ctx_solve_v2 (producer warps step, consumer warps evaluate + store, mbarrier ring) is the
    same arithmetic as ctx_solve_v1 on different warps: same bits, same status / step counts --
    for batches smaller than the grid, ragged T, failed rows (inf fill), constant states."""
    import sympy as sp

    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    if S == 3:
        model = cases.michaelis_menten()
        y0, theta, consts = cases.mm_inputs(B, seed=5, dtype=dtype)
        theta[:7, 0] = 1e-3                  # stiff rows -> max_steps -> inf fill
        t1 = 100.0
    else:
        states = [f"x{i}" for i in range(S)]
        params = [f"k{i}" for i in range(S)]
        odes = {states[i]: -sp.Symbol(params[i]) * sp.Symbol(states[i])
                + (sp.Float(0.5) * sp.Symbol(states[i - 1]) if i else 0) for i in range(S)}
        if S == 4:
            odes[states[3]] = sp.Integer(0)
        model = (states, params, [], odes, [])
        rng = np.random.default_rng(S)
        y0 = rng.uniform(0.5, 2.0, (B, S)).astype(dtype)
        theta = rng.uniform(0.05, 1.0, (B, S)).astype(dtype)
        consts = np.zeros((B, 0), dtype)
        t1 = 20.0
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    ts = np.linspace(0, t1, T).astype(dtype)
    outs = []
    for kernel in (2, 5):
        try:
            o = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), in_axes=(0, 0, 0, None), solver=solver,
                        rtol=1e-6, atol=1e-6, max_steps=512, kernel=kernel)
        except engine.EngineError as exc:
            # an explicit request that does not fit shared memory is an error (the automatic
            # selection falls back to the persistent kernel instead)
            assert kernel == 5 and "shared memory" in str(exc)
            pytest.skip("step-record ring of this model does not fit shared memory")
        torch.cuda.synchronize()
        outs.append([x.cpu().numpy() for x in (o.ys, o.status, o.n_accepted, o.n_rejected)])
    for a_, b_ in zip(*outs):
        np.testing.assert_array_equal(a_, b_)
    if S == 3:
        assert (outs[0][1] != 0).any()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("T,shared_ts,solver", [(16, True, "tsit5"), (7, False, "dopri5"), (32, True, "euler"),
                                                (1, True, "tsit5")])
def test_direct_save_variant_is_bit_identical(T, shared_ts, solver, dtype, monkeypatch):
    """Sparse grids (T <= 32) take the variant of ctx_solve_v1 whose lanes evaluate and store their
    own requested times (no cooperative pass): the same polynomial in the same Horner order, so the
    same bits as the cooperative path -- incl. failed rows (inf fill), t0 == t1 (constant fill),
    constant states and per-row grids."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    B = 4099
    model = cases.michaelis_menten()
    y0, theta, consts = cases.mm_inputs(B, seed=9, dtype=dtype)
    theta[:5, 0] = 1e-3                       # stiff rows -> max_steps -> inf fill
    rng = np.random.default_rng(2)
    if shared_ts:
        ts = np.linspace(0, 100, T).astype(dtype) if T > 1 else np.array([0.0], dtype=dtype)
    else:
        ts = np.sort(rng.uniform(0, 100, (B, T)), axis=1).astype(dtype)
        ts[:, 0] = 0.0
        ts[3] = 0.0                            # t0 == t1: every requested time is t0
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("CTX_B200_DIRECT_SAVE", flag)
        o = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), solver=solver,
                    in_axes=(0, 0, 0, None if shared_ts else 0), rtol=1e-6, atol=1e-6,
                    dt0=0.05 if solver == "euler" else 0.1, max_steps=60 if solver != "euler" else 4096,
                    kernel=2)
        torch.cuda.synchronize()
        assert engine.last_solve_kernel() == ("ctx_solve_v1 (direct save)" if flag == "1" else "ctx_solve_v1")
        res[flag] = (o.ys.cpu().numpy(), o.status.cpu().numpy(), o.n_accepted.cpu().numpy())
    np.testing.assert_array_equal(res["0"][0], res["1"][0])
    np.testing.assert_array_equal(res["0"][1], res["1"][1])
    np.testing.assert_array_equal(res["0"][2], res["1"][2])
    if solver != "euler" and T > 1:      # max_steps = 60: many rows fail -> inf fill in both variants
        failed = res["1"][1] == 1
        assert 10 < failed.sum() < B and np.isinf(res["1"][0][failed][:, -1, 1:]).all()
        assert np.isfinite(res["1"][0][~failed]).all()
    monkeypatch.delenv("CTX_B200_DIRECT_SAVE")
    # the automatic choice for T <= 32 is the direct variant
    m.solve(tt(y0), tt(theta), tt(consts), tt(ts), solver=solver, in_axes=(0, 0, 0, None if shared_ts else 0),
            dt0=0.05 if solver == "euler" else 0.1)
    assert engine.last_solve_kernel() == "ctx_solve_v1 (direct save)"


@pytest.mark.parametrize("kernel,T,B", [(2, 64, 5003), (2, 16, 5003), (5, 400, 70001)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_order_hint_changes_the_schedule_not_the_results(kernel, T, B, dtype):
    """``order=`` hands the rows to the persistent warps in a caller-given order (cost-ordered from
    a previous solve: the longest trajectories start first).  Any permutation must give the same
    bits, step counts and status for every row -- through the cooperative, the direct-save and the
    warp-specialised kernel."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    model = cases.michaelis_menten()
    y0, theta, consts = cases.mm_inputs(B, seed=21, dtype=dtype)
    theta[:9, 0] = 1e-3
    ts = np.linspace(0, 100, T).astype(dtype)
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    kw = dict(in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, max_steps=300, kernel=kernel)
    base = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), **kw)
    torch.cuda.synchronize()
    order = engine.cost_order(base.n_accepted, base.n_rejected, heavy_fraction=0.0625)
    # (the default of SolveOutput.cost_order(): the heaviest 1/16 first, the rest in index order)
    assert order.dtype == torch.int32 and sorted(order.tolist()) == list(range(B))
    steps = base.n_accepted + base.n_rejected
    k = round(B * 0.0625)
    head, rest = steps[order[:k].long()], order[k:]
    assert bool((head[:-1] >= head[1:]).all()) and int(head.min()) >= int(steps[rest.long()].max())
    assert bool((rest[1:] > rest[:-1]).all())
    full = engine.cost_order(base.n_accepted, base.n_rejected, heavy_fraction=1.0)
    cost = steps[full.long()]
    assert bool((cost[:-1] >= cost[1:]).all()) and sorted(full.tolist()) == list(range(B))
    rnd = torch.as_tensor(np.random.default_rng(0).permutation(B).astype(np.int32), device=dev)
    for o in (order, full, rnd):
        got = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), order=o, **kw)
        torch.cuda.synchronize()
        assert torch.equal(got.ys, base.ys) or bool(((got.ys == base.ys) | (got.ys.isnan() & base.ys.isnan())).all())
        assert torch.equal(got.status, base.status) and torch.equal(got.n_accepted, base.n_accepted)
        assert torch.equal(got.n_rejected, base.n_rejected)
    assert int((base.status != 0).sum()) > 0          # failed rows (inf fill) are part of the comparison
    with pytest.raises(ValueError, match="order must be"):
        m.solve(tt(y0), tt(theta), tt(consts), tt(ts), order=order[:-1].contiguous(), **kw)
    with pytest.raises(ValueError, match="order must be"):
        m.solve(tt(y0), tt(theta), tt(consts), tt(ts), order=order.long(), **kw)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kernel,T", [(1, 40), (2, 16), (2, 100), (5, 400)])
def test_zero_rhs_states_skip_the_error_norm_and_stage_sums_bit_identically(kernel, T, dtype, monkeypatch):
    """States whose generated right-hand side is identically zero (CTX_CONST_MASK: the enzyme of
    Michaelis-Menten) are left out of the error norm's division and of the stage sums.  That must
    not change one bit, one step count or one status -- including the cases where the skipped
    quotient would have been NaN: a constant state that is 0 with atol = 0 (0 / 0), NaN, or inf."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    model = cases.michaelis_menten()
    B = 20011
    y0, theta, consts = cases.mm_inputs(B, seed=13, dtype=dtype)
    y0[:5, 0] = 0.0              # E = 0: no reaction; with atol = 0 the reference's norm is 0 / 0 = NaN
    y0[5:8, 0] = np.nan
    y0[8:11, 0] = np.inf
    ts = np.linspace(0, 100, T).astype(dtype)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    for atol in (1e-6, 0.0):
        res = []
        for flags in ("", "-DCTX_NORM_SKIP_CONST=0 -DCTX_STAGE_SKIP_CONST=0"):
            monkeypatch.setenv("CTX_B200_NVRTC_FLAGS", flags)
            m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), np.dtype(dtype).name)
            o = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), in_axes=(0, 0, 0, None), rtol=1e-6,
                        atol=atol, max_steps=400, kernel=kernel)
            torch.cuda.synchronize()
            res.append(o)
        a, b = res
        assert torch.equal(a.status, b.status)
        assert torch.equal(a.n_accepted, b.n_accepted) and torch.equal(a.n_rejected, b.n_rejected)
        same = (a.ys == b.ys) | (a.ys.isnan() & b.ys.isnan())
        assert bool(same.all()), int((~same).sum())
        assert int((a.status[:11] != 0).sum()) >= 6         # the NaN / inf rows fail in both builds
        assert int((a.status[11:] == 0).sum()) > 0.9 * (B - 11)
