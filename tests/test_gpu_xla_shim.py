"""The XLA GPU custom-call shims (ctx_xla_solve / ctx_xla_solve_sens / ctx_xla_loglik_grad) driven
the way XLA's runtime drives a custom call: a ``void* buffers[]`` of operands then results, the
packed descriptor as opaque bytes, XLA's stream -- here a non-default torch stream -- and no
other argument (reference seam catalax/tools/simulation.py:160-196, called from
jit-compiled code at model.py:699 and mcmc.py:448).  Results must equal the direct C-ABI entry
points bit for bit; failures must be parked in the status slot."""
import ctypes as C

import numpy as np
import pytest

from catalax_b200 import workloads as wl

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _model(sens=False, dtype="float64"):
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    return engine.CompiledModel(cg.generate_field(wl.field_spec(wl.michaelis_menten()), sens_theta=sens), dtype)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_solve_and_sens_through_the_custom_call_abi(dtype):
    from catalax_b200 import engine

    m = _model(sens=True, dtype=dtype)
    dev = torch.device("cuda:0")
    y0, th, c = wl.mm_inputs(3000, seed=4, dtype=np.dtype(dtype))
    tt = lambda a: torch.as_tensor(a, device=dev)
    ts = torch.linspace(0, 100, 24, dtype=tt(y0).dtype, device=dev)
    kw = dict(in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6)
    direct = m.solve(tt(y0), tt(th), tt(c), ts, tangents=True, **kw)
    direct_plain = m.solve(tt(y0), tt(th), tt(c), ts, **kw)       # (another kernel build: no tangents)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                      # "XLA's" stream
        via = m.solve_xla(tt(y0), tt(th), tt(c), ts, tangents=True, **kw)
        plain = m.solve_xla(tt(y0), tt(th), tt(c), ts, **kw)
    side.synchronize()
    assert torch.equal(via.ys, direct.ys) and torch.equal(via.sens, direct.sens)
    assert torch.equal(via.n_accepted, direct.n_accepted) and torch.equal(via.status, direct.status)
    assert torch.equal(plain.ys, direct_plain.ys) and plain.sens is None
    assert torch.equal(plain.n_rejected, direct_plain.n_rejected)
    assert engine.lib().ctx_xla_last_status() == 0


def test_loglik_grad_through_the_custom_call_abi():
    from catalax_b200 import engine
    from oracle.c_oracle import COracle

    mm = wl.michaelis_menten()
    clean = lambda y0s, th, ts: COracle(*mm[:4], mm[4], dtype=np.float64).solve(
        y0s, np.tile(th, (len(y0s), 1)), np.zeros((len(y0s), 0)), ts, rtol=1e-8, atol=1e-8,
        max_steps=100000).ys
    y0s, theta, consts, ts, data, sigma = wl.nuts_operands(200, 12, 20, 2, clean)
    m = _model(sens=True)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)
    mask = torch.as_tensor(np.random.default_rng(0).uniform(size=data.shape) > 0.1, device=dev)
    for mk in (None, mask):
        d_out, d_part, d_st = m.loglik_grad(tt(y0s), tt(theta), tt(consts), tt(ts), tt(data), tt(sigma),
                                            mask=mk, likelihood="normal")
        x_out, x_part, x_st = m.loglik_grad_xla(tt(y0s), tt(theta), tt(consts), tt(ts), tt(data),
                                                tt(sigma), mask=mk, likelihood="normal")
        torch.cuda.synchronize()
        assert torch.equal(x_out, d_out) and torch.equal(x_part, d_part) and torch.equal(x_st, d_st)
    h_out, _, h_st = m.loglik_grad_host(y0s, theta, consts, ts, data, sigma, mask=mask.cpu().numpy(),
                                        likelihood="normal")
    assert np.array_equal(h_out, d_out.cpu().numpy()) and np.array_equal(h_st, d_st.cpu().numpy())
    assert engine.lib().ctx_xla_last_status() == 0


def test_custom_call_failures_are_parked_not_raised_inside_the_call():
    from catalax_b200 import engine

    m = _model()
    dev = torch.device("cuda:0")
    y0, th, c = wl.mm_inputs(8, seed=1, dtype=np.float64)
    tt = lambda a: torch.as_tensor(a, device=dev)
    ts = torch.linspace(0, 1, 4, dtype=torch.float64, device=dev)
    with pytest.raises(engine.EngineError, match="sensitivity directions|ctx_xla_solve_sens"):
        m.solve_xla(tt(y0), tt(th), tt(c), ts, tangents=True, in_axes=(0, 0, 0, None))   # no tangents generated
    assert engine.lib().ctx_xla_last_status() == 0        # the check consumed the parked status
    # workspace too small: the descriptor says so, the call refuses and parks CTX_E_WORKSPACE
    d = engine.ctx_xla_solve_desc()
    d.abi, d.model = engine.ABI_VERSION, m._h.value
    d.cfg = m.make_cfg(solver="tsit5", in_axes=(0, 0, 0, None), batch=8, n_times=4, max_steps=100,
                       rtol=1e-5, atol=1e-5, dt0=0.1)
    d.ws_bytes = 8
    ys = torch.empty((8, 4, 3), dtype=torch.float64, device=dev)
    cnt = [torch.empty(8, dtype=torch.int32, device=dev) for _ in range(3)]
    ws = torch.empty(8, dtype=torch.uint8, device=dev)
    bufs = [tt(y0), tt(th), tt(c), ts, ys] + cnt + [ws]
    arr = (C.c_void_p * len(bufs))(*[b.data_ptr() if b.numel() else 0 for b in bufs])
    raw = bytes(d)
    engine.lib().ctx_xla_solve(C.c_void_p(torch.cuda.current_stream().cuda_stream), arr, raw, len(raw))
    assert engine.lib().ctx_xla_last_status() == -5
