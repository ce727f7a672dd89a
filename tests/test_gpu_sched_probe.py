"""The library's opt-in scheduling probe (csrc/ctx_sched.cu): a capped solve of the same batch at a
looser tolerance that marks the rows still running at the cap, so that the real launch starts them
first.  It may change the schedule only -- never a bit of the result (same contract as the
caller's ``order=`` hint)."""
import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _same(a, b):
    return torch.equal(a, b) or bool(((a == b) | (a.isnan() & b.isnan())).all())


@pytest.mark.parametrize("kernel,T", [(0, 16), (2, 100), (5, 100)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_probe_changes_the_schedule_not_the_results(kernel, T, dtype):
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    B = (1 << 18) + 77
    model = cases.michaelis_menten()
    y0, theta, consts = cases.mm_inputs(B, seed=5, dtype=dtype)
    theta[:9, 0] = 1e-3                                   # rows that fail (inf fill) are compared too
    ts = np.linspace(0, 100, T).astype(dtype)
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), np.dtype(dtype).name)
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    kw = dict(in_axes=(0, 0, 0, None), rtol=1e-6, atol=1e-6, max_steps=300, kernel=kernel)
    n0 = engine.launch_count()
    base = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), **kw)
    torch.cuda.synchronize()
    n1 = engine.launch_count()
    got = m.solve(tt(y0), tt(theta), tt(consts), tt(ts), probe=True, **kw)
    torch.cuda.synchronize()
    n2 = engine.launch_count()
    assert n1 - n0 == 1
    assert n2 - n1 == 5, "probe solve + 3 partition launches + the solve itself"
    assert _same(got.ys, base.ys)
    assert torch.equal(got.status, base.status)
    assert torch.equal(got.n_accepted, base.n_accepted) and torch.equal(got.n_rejected, base.n_rejected)
    assert int((base.status != 0).sum()) > 0


def test_probe_is_opt_in_and_skipped_outside_its_window():
    """Off by default; with ``probe=True`` small batches and an explicit order still skip it."""
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg

    model = cases.michaelis_menten()
    m = engine.CompiledModel(cg.generate_field(cases.field_spec(model)), "float32")
    dev = torch.device("cuda:0")
    tt = lambda a: torch.as_tensor(a, device=dev)
    ts = tt(np.linspace(0, 100, 16).astype(np.float32))

    def launches(B, **kw):
        y0, theta, consts = cases.mm_inputs(B, seed=1, dtype=np.float32)
        n0 = engine.launch_count()
        out = m.solve(tt(y0), tt(theta), tt(consts), ts, in_axes=(0, 0, 0, None), max_steps=300,
                      rtol=1e-6, atol=1e-6, **kw)
        torch.cuda.synchronize()
        return engine.launch_count() - n0, out

    assert launches(1 << 16)[0] == 1
    assert launches(1 << 12, probe=True)[0] == 1
    n, out = launches(1 << 16, probe=True)
    assert n == 5
    assert launches(1 << 16, probe=True, order=out.cost_order())[0] == 1
