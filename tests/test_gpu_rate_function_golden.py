"""``Model._setup_rate_function`` / ``Model.rates`` / ``predict_rates`` (K4, ``ctx_rates``) on the GPU
against the values of the reference's REAL stacks (tests/golden/stack_reference_values.npz: what the
reference's ``_setup_rate_function(in_axes=None)`` returns IS its stack object)."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests.test_stack_golden import DEFS, G, _build

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(DEFS))
def test_rate_function_equals_the_reference_stack(name):
    import catalax_b200 as ctx

    ctx.enable_x64(True)
    try:
        m = _build(DEFS[name])
        f = m._setup_rate_function()
        t, y, th, c, y0, want = (G[f"{name}/{k}"] for k in ("t", "y", "theta", "consts", "y0", "dy"))
        got = f(t, y, (th, c, y0))                                  # batched: one launch
        scale = np.maximum(1.0, np.abs(want))
        assert (np.abs(got - want) <= 1e-12 * scale).all(), float((np.abs(got - want) / scale).max())
        one = f(float(t[3]), y[3], (th[3], c[3], y0[3]))            # the reference's per-point call form
        assert one.shape == want[3].shape and (np.abs(one - want[3]) <= 1e-12 * scale[3]).all()
    finally:
        ctx.enable_x64(False)


def test_model_rates_and_predict_rates():
    import catalax_b200 as ctx

    ctx.enable_x64(True)
    try:
        m = ctx.Model(name="MM")
        m.add_state(S="Substrate", P="Product")
        m.add_constant(E="Enzyme")
        m.add_ode("S", "-k_cat * E * S / (K_m + S)")
        m.add_ode("P", "k_cat * E * S / (K_m + S)")
        m.parameters.K_m.value, m.parameters.k_cat.value = 30.0, 0.4
        ds = ctx.Dataset.from_model(m)
        rng = np.random.default_rng(0)
        for i, e in enumerate((5.0, 12.0)):
            ds.add_measurement(ctx.Measurement(id=f"m{i}", initial_conditions={"S": 100.0, "P": 0.0, "E": e},
                                               time=np.linspace(0, 10, 4),
                                               data={"S": rng.uniform(10, 100, 4), "P": rng.uniform(0, 50, 4)}))
        r = m.predict_rates(ds)
        data, _, _ = ds.to_jax_arrays(m.get_state_order())
        pts = np.asarray(data).reshape(8, 2)                       # state order [P, S]
        E = np.repeat([5.0, 12.0], 4)
        v = 0.4 * E * pts[:, 1] / (30.0 + pts[:, 1])
        np.testing.assert_allclose(r, np.stack([v, -v], axis=1), rtol=1e-12)
        with pytest.raises(AssertionError, match="same number of rows"):
            m.rates(np.zeros(3), pts)
    finally:
        ctx.enable_x64(False)
