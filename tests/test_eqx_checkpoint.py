"""The reference's trained NeuralODE checkpoint (examples/trained/menten_trained.eqx, copied to
tests/golden/reference_fixtures/) pins the restated MLP conventions (mlp.py:71-84, eqx.nn.MLP):
loaded with those conventions, the network reproduces the Michaelis-Menten law it was trained on
(examples/models/menten_model.json: ds1/dt = -v_max*s1/(K_m+s1), v_max=7, K_m=100)."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).parent / "golden" / "reference_fixtures"


def _mm_rate(s):
    return -7.0 * s / (100.0 + s)


def test_checkpoint_loads_and_oracle_forward_matches_the_trained_law():
    from catalax_b200.neural import NeuralODE
    from oracle import mlp_oracle as mo

    net = NeuralODE.from_eqx(GOLD / "menten_trained.eqx")
    assert net.state_order == ["s1"] and net.spec.activation == "celu"
    assert [w.shape for w in net.weights] == [(4, 2), (4, 4), (1, 4)]
    assert net.biases[-1] is None and abs(net.max_time - 93.42408) < 1e-4
    rhs = mo.make_rhs("neuralode", net.weights, net.biases, "celu", "identity", net.max_time)
    s = np.linspace(20.0, 300.0, 15)[:, None]
    got = rhs(np.zeros(15), s, None, None, None)[:, 0]
    want = _mm_rate(s[:, 0])
    # a wrong convention (transposed weights, missing t/max_time input, other activation) is off
    # by O(1); the trained net is within its training error
    assert np.abs(got - want).max() < 0.4 and np.abs(got / want - 1).mean() < 0.05


@pytest.mark.gpu
def test_checkpoint_predicts_the_substrate_decay_on_the_gpu():
    import catalax_b200 as ctx
    from catalax_b200.neural import NeuralODE
    from oracle import diffrax_oracle as do, mlp_oracle as mo
    from tests import cases

    ctx.enable_x64(True)
    try:
        net = NeuralODE.from_eqx(GOLD / "menten_trained.eqx")
        ts = np.linspace(0.0, 90.0, 19)
        y0s = np.array([[300.0], [200.0], [100.0], [50.0]])
        ys = net.predict_arrays(ts, y0s)
        rhs = mo.make_rhs("neuralode", net.weights, net.biases, "celu", "identity", net.max_time)
        ref = do.solve(rhs, y0s, np.zeros(0), np.zeros((4, 0)), ts, rtol=1e-3, atol=1e-6,
                       dt0=float(ts[1] - ts[0]), dtype=np.float64, t0_from_ts=True)
        np.testing.assert_allclose(ys, ref.ys, rtol=1e-6, atol=1e-9)
        for row, s0 in enumerate(y0s[:, 0]):
            exact = cases.mm_closed_form(ts, s0, 7.0, 100.0)
            assert np.abs(ys[row, :, 0] - exact).max() < 0.06 * s0      # training error level
        got = net.rates(np.zeros(4), y0s)
        np.testing.assert_allclose(got[:, 0], _mm_rate(y0s[:, 0]), rtol=0.12)
    finally:
        ctx.enable_x64(False)
