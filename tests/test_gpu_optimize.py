"""The `optimize` residual loop (catalax/tools/optimization.py:14-207) on the GPU:
every objective evaluation is one batched solve with forward tangents (ctx_solve_sens)."""
import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

THETA_TRUE = {"K_m": 40.0, "k_cat": 0.3}


def _setup(noise=0.0, seed=0):
    import catalax_b200 as ctx
    from oracle import diffrax_oracle as do, symbolic as sy

    model = ctx.Model(name="MM")
    model.add_state(S="Substrate", E="Enzyme", P="Product")
    model.add_ode("S", "-k_cat * E * S / (K_m + S)")
    model.add_ode("P", "k_cat * E * S / (K_m + S)")
    model.add_ode("E", "0")
    model.parameters.K_m.initial_value = 70.0
    model.parameters.K_m.lower_bound, model.parameters.K_m.upper_bound = 1.0, 500.0
    model.parameters.k_cat.initial_value = 0.15
    rng = np.random.default_rng(seed)
    M, T = 6, 12
    order = model.get_state_order()                      # [E, P, S]
    y0s = np.stack([rng.uniform(5, 20, M), np.zeros(M), rng.uniform(50, 300, M)], 1)
    times = np.tile(np.linspace(0, 80, T), (M, 1))
    mm = cases.michaelis_menten()
    rhs, *_ = sy.make_rhs(*mm[:4], mm[4])
    clean = do.solve(rhs, y0s, np.array([THETA_TRUE["K_m"], THETA_TRUE["k_cat"]]), np.zeros((M, 0)),
                     times, rtol=1e-10, atol=1e-10).ys
    data = clean + noise * rng.normal(size=clean.shape)
    ds = ctx.Dataset.from_jax_arrays(order, data, times, y0s)
    return model, ds


@pytest.mark.parametrize("method,objective,tol", [
    ("leastsq", "plain", 5e-4),       # residual = pred - data: Gauss-Newton (data: rtol 1e-10, fit: 1e-5)
    ("bfgs", "l2", 5e-2),             # the reference's defaults: sum((0.5 e^2)^2), flat at the optimum
])
def test_optimize_recovers_the_parameters(method, objective, tol):
    import catalax_b200 as ctx
    from catalax_b200.tools.optimization import l2_loss

    ctx.enable_x64(True)
    try:
        model, ds = _setup()
        fun = l2_loss if objective == "l2" else (lambda pred, target: pred - target)
        # (global upper bound 5: k_cat = 1e5 with K_m = 1 is a stiffness ratio no explicit
        # solver integrates in less than millions of steps)
        result, fitted = ctx.optimize(model, ds, dt0=0.01, objective_fun=fun, method=method,
                                      global_upper_bound=5.0)
        for name, true in THETA_TRUE.items():
            assert abs(result.params[name] - true) <= tol * true, (name, result.params, result.message)
            assert fitted.parameters[name].value == result.params[name]
        assert model.parameters["K_m"].value is None          # the input model is not modified
        assert result.nfev > 2 and result.chisqr < 1e-2
        # bounds: the global bounds were filled in for the parameter that had none
        assert model.parameters["k_cat"].lower_bound == 1e-6 and model.parameters["k_cat"].upper_bound == 5.0
    finally:
        ctx.enable_x64(False)


def test_objective_gradient_is_the_exact_gradient_of_the_discrete_solve():
    import catalax_b200 as ctx
    from catalax_b200.tools.optimization import _Objective, l2_loss

    ctx.enable_x64(True)
    try:
        model, ds = _setup(noise=0.5, seed=3)
        obj = _Objective(model, ds, 0.01, 64 ** 4, l2_loss)
        x = np.array([55.0, 0.25])
        F, g, r, J = obj(x)
        assert r.shape == (6 * 12 * 3,) and J.shape == (6 * 12 * 3, 2)
        np.testing.assert_allclose(g, 2.0 * (r[:, None] * J).sum(0), rtol=1e-12)
        # The gradient differentiates the accepted step sequence (the controller is under
        # stop_gradient in the reference); a finite difference of the adaptive solve also sees
        # O(rtol) jumps where accept/reject decisions flip, hence the wide step and tolerance.
        for j in range(2):
            h = 1e-3 * x[j]
            xp, xm = x.copy(), x.copy()
            xp[j] += h; xm[j] -= h
            fd = (obj(xp)[0] - obj(xm)[0]) / (2 * h)
            assert abs(g[j] - fd) <= 5e-3 * abs(fd), (j, g[j], fd)
    finally:
        ctx.enable_x64(False)


def test_missing_initial_value_is_an_error():
    import catalax_b200 as ctx

    model, ds = _setup()
    model.parameters.k_cat.initial_value = None
    with pytest.raises(ValueError, match="initial"):
        ctx.optimize(model, ds)
