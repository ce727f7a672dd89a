"""The oracle's solver coefficients against order conditions and scipy."""
import numpy as np

from oracle import tableaus as tb


def _A(rows, n):
    A = np.zeros((n, n))
    for i, r in enumerate(rows):
        A[i, : len(r)] = r
    return A


def _order_conditions(A, b, c, order):
    e = np.ones_like(c)
    conds = [(b @ e, 1.0), (b @ c, 1 / 2)]
    if order >= 3:
        conds += [(b @ c**2, 1 / 3), (b @ (A @ c), 1 / 6)]
    if order >= 4:
        conds += [(b @ c**3, 1 / 4), (b @ (c * (A @ c)), 1 / 8), (b @ (A @ c**2), 1 / 12),
                  (b @ (A @ (A @ c)), 1 / 24)]
    if order >= 5:
        Ac, Ac2, AAc = A @ c, A @ c**2, A @ (A @ c)
        conds += [(b @ c**4, 1 / 5), (b @ (c**2 * Ac), 1 / 10), (b @ (Ac * Ac), 1 / 20),
                  (b @ (c * Ac2), 1 / 15), (b @ (A @ c**3), 1 / 20), (b @ (c * AAc), 1 / 30),
                  (b @ (A @ (c * Ac)), 1 / 40), (b @ (A @ Ac2), 1 / 60), (b @ (A @ AAc), 1 / 120)]
    return max(abs(l - r) for l, r in conds)


def test_tsit5_order_conditions():
    A, c = _A(tb.TSIT5_A, 7), tb.TSIT5_C
    assert np.abs(A.sum(1) - c).max() < 1e-15
    assert _order_conditions(A, tb.TSIT5_B, c, 5) < 5e-15
    assert _order_conditions(A, tb.TSIT5_BHAT, c, 4) < 5e-15
    assert _order_conditions(A, tb.TSIT5_BHAT, c, 5) > 1e-5  # embedded weights are 4th order
    assert abs(tb.TSIT5_E[-1] + 1 / 66) < 1e-17


def test_tsit5_dense_output_is_consistent():
    A, c, b = _A(tb.TSIT5_A, 7), tb.TSIT5_C, tb.TSIT5_B
    assert np.abs(np.array(tb.tsit5_dense_weights(0.0))).max() == 0.0
    assert np.abs(np.array(tb.tsit5_dense_weights(1.0)) - b).max() < 5e-15
    for th in (0.25, 0.5, 0.75):
        w = np.array(tb.tsit5_dense_weights(th))
        assert abs(w.sum() - th) < 5e-15
        assert abs(w @ c - th**2 / 2) < 5e-15
        assert abs(w @ c**2 - th**3 / 3) < 5e-15
        assert abs(w @ (A @ c) - th**3 / 6) < 5e-15
        assert abs(w @ c**3 - th**4 / 4) < 5e-15


def test_dopri5_matches_scipy_rk45():
    from scipy.integrate._ivp.rk import RK45

    A = _A(tb.DOPRI5_A, 7)
    assert np.abs(A[:6, :5] - RK45.A[:6, :5]).max() < 1e-16
    assert np.abs(tb.DOPRI5_B[:6] - RK45.B).max() < 1e-16
    assert np.abs(tb.DOPRI5_C[:6] - RK45.C).max() < 1e-16
    assert np.abs(tb.DOPRI5_E + RK45.E).max() < 1e-16  # scipy stores b_hat - b
    assert _order_conditions(A, tb.DOPRI5_B, tb.DOPRI5_C, 5) < 5e-16
    assert _order_conditions(A, tb.DOPRI5_BHAT, tb.DOPRI5_C, 4) < 5e-16


def test_dopri5_midpoint_weights_are_fourth_order():
    A, c, m = _A(tb.DOPRI5_A, 7), tb.DOPRI5_C, tb.DOPRI5_CMID
    th = 0.5
    assert abs(m.sum() - th) < 1e-15
    assert abs(m @ c - th**2 / 2) < 1e-15
    assert abs(m @ c**2 - th**3 / 3) < 1e-15
    assert abs(m @ (A @ c) - th**3 / 6) < 1e-15
    assert abs(m @ c**3 - th**4 / 4) < 2e-4  # Shampine's weights: 4th-order accurate
