"""Pin the oracle against every golden vector the reference's tests hold for this path
 and against closed forms.  CPU only.

Reference test -> here:
  tests/integration/test_simulation.py:41-69   exponential growth with a constant
  tests/integration/test_simulation.py:81-143  "{state}_init" symbols
  tests/integration/test_simulation.py:145-195 "_init" declared as constant + assignment
  tests/fixtures/data.npy / times.npy          f32 Michaelis-Menten trajectory (Lambert-W)
  tests/unit/simulation/test_stack.py:13-203   ODE / reaction / mixed stack algebra
"""
from pathlib import Path

import numpy as np
import pytest
import sympy as sp
from scipy.special import lambertw

from oracle import diffrax_oracle as do, symbolic as sy

TOLERANCE = 1e-3  # the reference's own bound (test_simulation.py:5)
GOLD = Path(__file__).parent / "golden" / "reference_fixtures"


GOLD_EXP = [np.array([1.0, 4.481674, 20.08544]),      # the reference's own f32 outputs
            np.array([1.0, 20.085447, 403.42603])]    # (test_simulation.py:52-63); e^6 = 403.42879
FIRST_FACTORS = np.round(np.arange(1.0, 10.0001, 0.01), 2)


def fit_first_factor(solve_rows, row):
    """min over the step-size factor that follows loop iteration 0 of sum|y - golden| for one row
    of the exponential-growth golden; ``solve_rows(e, factors) -> ys [len(factors), 3]``."""
    ys = np.asarray(solve_rows([1.0, 2.0][row], FIRST_FACTORS), dtype=np.float64)
    res = np.abs(ys - GOLD_EXP[row]).sum(axis=1)
    k = int(np.argmin(res))
    return float(FIRST_FACTORS[k]), float(res[k]), ys[k]


def _oracle_rows(**cfg):
    rhs, *_ = sy.make_rhs(["s1"], ["q"], ["e"], {"s1": "e * s1 * q"})
    ts = np.linspace(0, 3, 3).astype(np.float32)

    def solve_rows(e, factors):
        B = len(factors)
        fo = np.full((B, 1), np.nan)
        fo[:, 0] = factors
        kw = {**dict(dtype=np.float32), **cfg}
        return do.solve(rhs, np.ones((B, 1)), [1.0], np.full((B, 1), e), ts, factor_override=fo,
                        **kw).ys[:, :, 0]
    return solve_rows


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_exponential_growth_with_constant(dtype):
    """tests/integration/test_simulation.py:41-69 with the reference's own bound where it is
    implementation-independent (row 1), and the solver-tolerance band for row 2 -- whose golden
    value is ONE f32 step path of diffrax, pinned by the fitted test below."""
    rhs, *_ = sy.make_rhs(["s1"], ["q"], ["e"], {"s1": "e * s1 * q"})
    ts = np.linspace(0, 3, 3).astype(dtype)
    assert ts.tolist() == [0.0, 1.5, 3.0]
    r = do.solve(rhs, [[1.0], [1.0]], [1.0], [[1.0], [2.0]], ts, dtype=dtype)  # default config
    assert np.abs(r.ys[0, :, 0] - GOLD_EXP[0]).sum() < TOLERANCE
    assert (np.abs(r.ys[1, :, 0] / GOLD_EXP[1] - 1) < 2e-5).all()
    assert (np.abs(r.ys[1, :, 0] / np.exp(2 * ts.astype(np.float64)) - 1) < 2e-5).all()


def test_exponential_growth_golden_is_reproduced_up_to_the_first_step_factor():
    """The golden 403.42603 differs from e^6 by 2.8e-3: it is the discretisation error of the
    reference's accepted steps.  The first step [0, 0.1] has a true error estimate ~1e-9
    relative, i.e. in f32 its value -- and with it the factor (anywhere in [1, 10]) the
    accept-never-shrinks controller multiplies the step size by -- is rounding noise that no
    second implementation reproduces.  With that ONE factor as a nuisance parameter the oracle
    (Tsit5, rtol = atol = 1e-5, f32: the reference's default config) reproduces BOTH goldens to
    f32 rounding with the reference's own bound to spare (two numbers per row, one parameter):
    row 1 at factor 4.15 (sum|err| 2.0e-6, ulp(20) = 1.9e-6), row 2 at 4.41 (1.6e-5,
    ulp(403) = 3.1e-5).  No other solver or tolerance gets within 5x of those bounds (measured:
    Dopri5 1.6e-4 / 6.2e-3, rtol x2 4.4e-5 / 4.2e-3, rtol / 2 1.7e-5 / 1.4e-4, f64 arithmetic
    1.7e-5 / 3.8e-4)."""
    bounds = (2.5e-6, 2.5e-5)
    for row in (0, 1):
        f, res, ys = fit_first_factor(_oracle_rows(), row)
        assert res < bounds[row] < TOLERANCE, (row, f, res)
        assert 1.0 < f < 10.0          # an interior minimum, not the edge of the scan
    for cfg in (dict(solver="dopri5"), dict(rtol=2e-5, atol=2e-5), dict(rtol=5e-6, atol=5e-6),
                dict(dtype=np.float64)):
        for row in (0, 1):
            f, res, ys = fit_first_factor(_oracle_rows(**cfg), row)
            assert res > 5 * bounds[row], (cfg, row, f, res)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_init_symbol(dtype):
    rhs, *_ = sy.make_rhs(["p", "s"], ["k"], [], {"s": "-k * s", "p": "s_init"})
    ts = np.linspace(0, 3, 4).astype(dtype)
    r = do.solve(rhs, [[0.0, 2.0], [0.0, 5.0]], [1.0], np.zeros((2, 0)), ts, dtype=dtype)
    t = np.array([0.0, 1.0, 2.0, 3.0])
    assert np.abs(r.ys[0, :, 0] - 2.0 * t).sum() < TOLERANCE
    assert np.abs(r.ys[1, :, 0] - 5.0 * t).sum() < TOLERANCE
    assert np.abs(r.ys[0, :, 1] - 2.0 * np.exp(-t)).sum() < TOLERANCE


def test_init_symbol_via_assignment():
    # rate = k * s_init inlined (model.py:788-820)
    rhs, *_ = sy.make_rhs(["p", "s"], ["k"], [], {"s": "-(k * s_init)", "p": "k * s_init"})
    ts = np.linspace(0, 2, 4)
    r = do.solve(rhs, [[0.0, 2.0], [0.0, 4.0]], [0.5], np.zeros((2, 0)), ts, dtype=np.float32)
    for i, s_init in enumerate([2.0, 4.0]):
        assert np.abs(r.ys[i, :, 0] - 0.5 * s_init * ts).sum() < TOLERANCE


def _mm_closed_form(t, s0, V, K):
    return K * np.real(lambertw(s0 / K * np.exp((s0 - V * t) / K)))


def test_reference_fixture_trajectory():
    """tests/fixtures/data.npy is a real f32 output of the reference for
    ds/dt = -v_max*s/(K_m+s), v_max=10, K_m=100, s0=100 on linspace(0,100,100).
    It agrees with the closed form to f32 rounding, i.e. it was produced with a tight
    tolerance; the oracle must converge onto it as the tolerance is tightened."""
    data = np.load(GOLD / "data.npy")[0, :, 0].astype(np.float64)
    t = np.load(GOLD / "times.npy").astype(np.float64)
    rhs, *_ = sy.make_rhs(["s1"], ["K_m", "v_max"], [], {"s1": "-v_max*s1/(K_m+s1)"})
    exact = _mm_closed_form(t, 100.0, 10.0, 100.0)
    assert np.abs(data - exact).max() < 1.1e-5  # fixture == closed form up to f32 rounding
    prev = None
    for tol in (1e-5, 1e-7, 1e-9):
        r = do.solve(rhs, [[100.0]], [100.0, 10.0], np.zeros((1, 0)), t, rtol=tol, atol=tol,
                     dtype=np.float64)
        err = np.abs(r.ys[0, :, 0] - data).max()
        assert prev is None or err <= prev
        prev = err
    assert prev < 1.1e-5
    # default reference config in f32: within the reference's own TOLERANCE per point
    r32 = do.solve(rhs, [[100.0]], [100.0, 10.0], np.zeros((1, 0)), t.astype(np.float32),
                   dtype=np.float32)
    assert np.abs(r32.ys[0, :, 0] - data).max() < TOLERANCE


@pytest.mark.parametrize("solver", ["tsit5", "dopri5"])
def test_michaelis_menten_closed_form_and_sensitivities(solver):
    """README model: S(t) = K W((S0/K) exp((S0 - V t)/K)), V = k_cat E; closed-form
    sensitivities dS/dV = -t S/(K+S), dS/dK = S/K - S/(K+S) (1 + (S0 - V t)/K)."""
    from tests import cases

    m = cases.michaelis_menten()
    rhs, N, D, init_aug = sy.make_rhs(*m[:4], m[4], sens_theta=True)
    E0, S0, K, kc = 10.0, 100.0, 40.0, 0.3
    t = np.linspace(0, 60, 25)
    r = do.solve(rhs, [[E0, 0.0, S0]], [K, kc], np.zeros((1, 0)), t, solver=solver, rtol=1e-9,
                 atol=1e-9, dtype=np.float64, n_err=3, init_aug=init_aug(1, np.float64))
    V = kc * E0
    S = _mm_closed_form(t, S0, V, K)
    np.testing.assert_allclose(r.ys[0, :, 2], S, atol=5e-7)
    np.testing.assert_allclose(r.ys[0, :, 1], S0 - S, atol=5e-7)
    dS_dV = -t * S / (K + S)
    dS_dK = S / K - S / (K + S) * (1 + (S0 - V * t) / K)
    sens = r.ys[0, :, 3:].reshape(len(t), 2, 3)  # [T, dir, state]
    np.testing.assert_allclose(sens[:, 0, 2], dS_dK, atol=2e-6)
    np.testing.assert_allclose(sens[:, 1, 2], E0 * dS_dV, atol=2e-5)


def test_tangents_equal_finite_differences_of_the_frozen_step_sequence():
    """diffrax differentiates the discrete scheme with the step-size factor under
    stop_gradient: tangents must equal d/dtheta of a replay with the SAME steps."""
    from tests import cases

    m = cases.nonautonomous()
    rhs_aug, N, D, init_aug = sy.make_rhs(*m[:4], m[4], sens_theta=True, sens_y0=True)
    rhs, *_ = sy.make_rhs(*m[:4], m[4])
    rng = np.random.default_rng(3)
    y0 = rng.uniform(0.5, 1.5, (1, 4)); th = rng.uniform(0.2, 0.9, (1, 3)); c = np.array([[0.3]])
    ts = np.linspace(0, 4, 9)
    kw = dict(rtol=1e-6, atol=1e-8, dtype=np.float64)
    base = do.solve(rhs_aug, y0, th, c, ts, n_err=4, init_aug=init_aug(1, np.float64),
                    record_steps=True, **kw)
    plain = do.solve(rhs, y0, th, c, ts, **kw)
    np.testing.assert_array_equal(plain.n_accepted, base.n_accepted)  # tangents do not steer
    np.testing.assert_allclose(plain.ys, base.ys[..., :4], rtol=0, atol=1e-13)
    sens = base.ys[0, :, 4:].reshape(len(ts), D, 4)
    eps = 1e-6
    for d in range(D):
        def run(sign):
            y0p, thp = y0.copy(), th.copy()
            if d < 3:
                thp[0, d] += sign * eps
            else:
                y0p[0, d - 3] += sign * eps
            return do.solve(rhs, y0p, thp, c, ts, forced_steps=base.steps, **kw).ys[0]
        fd = (run(+1) - run(-1)) / (2 * eps)
        np.testing.assert_allclose(sens[:, d, :], fd, rtol=2e-6, atol=2e-8)


def test_stack_algebra_golden():
    """tests/unit/simulation/test_stack.py: ODE, reaction and mixed stacks at one point."""
    one = np.ones(1)
    k1, k2, k3, s1, s2 = 1.0, 2.0, 3.0, 1.0, 2.0
    Y = np.array([[s1, s2]])
    r1 = ("r1", [(1.0, "s1")], [(1.0, "s2")], "k1 * s1")
    r2 = ("r2", [(1.0, "s2")], [(1.0, "s1")], "k2 * s2")
    rhs, *_ = sy.make_rhs(["s1", "s2"], ["k1", "k2", "k3"], [], {"s1": "k3"}, [r1, r2])
    out = rhs(one, Y, np.array([[k1, k2, k3]]), np.zeros((1, 0)), Y)
    np.testing.assert_allclose(out[0], [-k1 * s1 + k2 * s2 + k3, k1 * s1 - k2 * s2])
    rhs, *_ = sy.make_rhs(["s1", "s2"], ["k1", "k2"], [], {"s1": "k1 * s1", "s2": "k2 * s2"})
    out = rhs(one, Y, np.array([[k1, k2]]), np.zeros((1, 0)), Y)
    np.testing.assert_allclose(out[0], [k1 * s1, k2 * s2])
    rhs, *_ = sy.make_rhs(["s1", "s2"], ["k1", "k2"], [], {}, [r1, r2])
    out = rhs(one, Y, np.array([[k1, k2]]), np.zeros((1, 0)), Y)
    np.testing.assert_allclose(out[0], [-k1 * s1 + k2 * s2, k1 * s1 - k2 * s2])


def test_throw_false_semantics_inf_after_failure():
    rhs, *_ = sy.make_rhs(["s1"], ["K_m", "v_max"], [], {"s1": "-v_max*s1/(K_m+s1)"})
    t = np.linspace(0, 100, 50)
    r = do.solve(rhs, [[100.0]], [100.0, 10.0], np.zeros((1, 0)), t, rtol=1e-10, atol=1e-10,
                 max_steps=30)
    assert r.status[0] == 1 and r.n_accepted[0] + r.n_rejected[0] == 30
    assert np.isfinite(r.ys[0, 0, 0]) and np.isinf(r.ys[0, -1, 0])
    k = np.isinf(r.ys[0, :, 0]).argmax()
    assert np.isinf(r.ys[0, k:, 0]).all() and np.isfinite(r.ys[0, :k, 0]).all()
