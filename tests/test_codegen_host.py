"""The CUDA code generator, checked on the CPU: the generated field is compiled as host C++
and compared with the oracle's independent sympy.lambdify evaluation (reference semantics of
catalax/tools/stack.py + symbolicmodule.py)."""
import numpy as np
import pytest
import sympy as sp

from catalax_b200.tools import codegen as cg
from oracle import symbolic as sy
from tests import cases
from tests.host_field import HostField

MODELS = {
    "mm": cases.michaelis_menten,
    "menten_fixture": cases.menten_fixture,
    "nonautonomous": cases.nonautonomous,
    "mass_action_8": lambda: cases.mass_action_network(8),
    "mass_action_32": lambda: cases.mass_action_network(32),
}


@pytest.mark.parametrize("name", list(MODELS))
def test_generated_rhs_matches_lambdify(name):
    model = MODELS[name]()
    states, params, consts, odes, reactions = model
    field = cg.generate_field(cases.field_spec(model))
    hf = HostField(field)
    rhs, *_ = sy.make_rhs(states, params, consts, odes, reactions)
    rng = np.random.default_rng(0)
    for _ in range(5):
        y = rng.uniform(0.1, 2.0, len(states)); th = rng.uniform(0.1, 1.5, len(params))
        c = rng.uniform(0.1, 1.0, len(consts)); y0 = rng.uniform(0.1, 2.0, len(states))
        t = rng.uniform(0, 3)
        want = rhs(np.array([t]), y[None], th[None], c[None], y0[None])[0]
        got = hf.rhs(t, y, th, c, y0)
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", ["mm", "nonautonomous", "mass_action_8"])
def test_generated_tangent_field_matches_lambdify(name):
    model = MODELS[name]()
    states, params, consts, odes, reactions = model
    S = len(states)
    field = cg.generate_field(cases.field_spec(model), sens_theta=True, sens_y0=True)
    assert field.D == len(params) + S and field.n_theta_dirs == len(params)
    hf = HostField(field)
    rhs, N, D, _ = sy.make_rhs(states, params, consts, odes, reactions, sens_theta=True, sens_y0=True)
    rng = np.random.default_rng(1)
    y = rng.uniform(0.1, 2.0, S); th = rng.uniform(0.1, 1.5, len(params))
    c = rng.uniform(0.1, 1.0, len(consts)); y0 = rng.uniform(0.1, 2.0, S)
    s = rng.normal(size=S * D)
    Y = np.concatenate([y, s])
    want = rhs(np.array([0.7]), Y[None], th[None], c[None], y0[None])[0]
    dy, ds = hf.tan(0.7, y, s, th, c, y0)
    np.testing.assert_allclose(dy, want[:S], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(ds, want[S:], rtol=1e-12, atol=1e-14)


def test_stack_golden_values_through_the_generator():
    """tests/unit/simulation/test_stack.py of the reference, through the CUDA generator."""
    k1, k2, k3, s1, s2 = 1.0, 2.0, 3.0, 1.0, 2.0
    r1 = ("r1", [(1.0, "s1")], [(1.0, "s2")], sp.sympify("k1 * s1"))
    r2 = ("r2", [(1.0, "s2")], [(1.0, "s1")], sp.sympify("k2 * s2"))
    f = cg.generate_field(cg.FieldSpec(["s1", "s2"], ["k1", "k2", "k3"], [], {"s1": sp.Symbol("k3")}, [r1, r2]))
    np.testing.assert_allclose(HostField(f).rhs(1.0, [s1, s2], [k1, k2, k3], [], [s1, s2]),
                               [-k1 * s1 + k2 * s2 + k3, k1 * s1 - k2 * s2])
    f = cg.generate_field(cg.FieldSpec(["s1", "s2"], ["k1", "k2"], [], {}, [r1, r2]))
    np.testing.assert_allclose(HostField(f).rhs(1.0, [s1, s2], [k1, k2], [], [s1, s2]),
                               [-k1 * s1 + k2 * s2, k1 * s1 - k2 * s2])


def test_operator_table_coverage():
    """Every op of symbolicmodule.py:52-98 that makes sense for a scalar field prints."""
    x, y, p = sp.symbols("x y p")
    exprs = [sp.Abs(x - y), sp.sign(x - 1), sp.ceiling(x), sp.floor(y), sp.log(x + 1), sp.exp(-x),
             sp.sqrt(x + y), sp.cos(x), sp.acos(x / 4), sp.sin(x), sp.asin(x / 4), sp.tan(x / 4),
             sp.atan(x), sp.atan2(x, y), sp.cosh(x / 3), sp.acosh(x + 1), sp.sinh(x / 3),
             sp.asinh(x), sp.tanh(x), sp.atanh(x / 4), x ** p, x ** -2, x ** sp.Rational(3, 2),
             sp.erf(x), sp.Max(x, y, p), sp.Min(x, y), sp.Rational(2, 3) * x, sp.pi * x]
    spec = cg.FieldSpec([f"s{i}" for i in range(len(exprs))], ["p"], [], {}, [])
    names = {sp.Symbol("x"): sp.Symbol("s0"), sp.Symbol("y"): sp.Symbol("s1")}
    spec.odes = {f"s{i}": e.xreplace(names) for i, e in enumerate(exprs)}
    field = cg.generate_field(spec)
    hf = HostField(field)
    yv = np.linspace(0.6, 1.9, len(exprs)); th = np.array([1.3])
    got = hf.rhs(0.0, yv, th, [], yv)
    for i, e in enumerate(exprs):
        want = float(e.xreplace(names).subs({sp.Symbol("s0"): yv[0], sp.Symbol("s1"): yv[1], p: th[0]}).evalf())
        assert abs(got[i] - want) <= 1e-12 * max(1, abs(want)), (e, got[i], want)


def test_missing_symbol_raises_like_the_reference():
    with pytest.raises(KeyError, match="Missing input for symbol"):
        cg.generate_field(cg.FieldSpec(["S"], ["k"], [], {"S": -sp.Symbol("k") * sp.Symbol("E") * sp.Symbol("S")}, []))


WARP_MODELS = {
    "nonautonomous": cases.nonautonomous,          # 4 classes, t, constants, *_init
    "mm": cases.michaelis_menten,                  # ODE form, a zero right-hand side
    "mass_action_32": lambda: cases.mass_action_network(32),   # 2 classes x 32 terms
    "mass_action_40": lambda: cases.mass_action_network(40),   # ragged second state slot
}


@pytest.mark.parametrize("name", list(WARP_MODELS))
def test_lane_parallel_form_equals_the_thread_form(name):
    """generate_warp_form (the field of the warp-per-trajectory kernel ctx_solve_w), emulated
    lane by lane on the host, equals ctx_rhs and the oracle's lambdify evaluation."""
    model = WARP_MODELS[name]()
    states, params, consts, odes, reactions = model
    field = cg.generate_field(cases.field_spec(model), warp_form=True)
    assert field.warp_form and field.warp_info["state_slots"] == (len(states) + 31) // 32
    hf = HostField(field)
    rhs, *_ = sy.make_rhs(states, params, consts, odes, reactions)
    rng = np.random.default_rng(4)
    for _ in range(4):
        y = rng.uniform(0.1, 2.0, len(states)); th = rng.uniform(0.1, 1.5, len(params))
        c = rng.uniform(0.1, 1.0, len(consts)); y0 = rng.uniform(0.1, 2.0, len(states))
        t = rng.uniform(0, 3)
        want = rhs(np.array([t]), y[None], th[None], c[None], y0[None])[0]
        np.testing.assert_allclose(hf.warp(t, y, th, c, y0), want, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(hf.warp(t, y, th, c, y0), hf.rhs(t, y, th, c, y0), rtol=1e-13, atol=1e-15)


def test_lane_parallel_form_is_automatic_only_for_large_structured_models():
    small = cg.generate_field(cases.field_spec(cases.michaelis_menten()))
    assert not small.warp_form and "CTX_HAVE_WARP" not in small.source
    big = cg.generate_field(cases.field_spec(cases.mass_action_network(32)))
    assert big.warp_form and big.warp_info["classes"] == 2 and big.warp_info["term_slots"] == 2
    # 40 unrelated rate laws: too many structural classes -> thread form only
    n = 40
    states = [f"s{i:02d}" for i in range(n)]
    odes = {s: sum(sp.Symbol(states[(i + j) % n]) ** (j + 1) for j in range(i % 13 + 1)) * sp.Symbol("k")
            for i, s in enumerate(states)}
    odd = cg.generate_field(cg.FieldSpec(states, ["k"], [], odes, []))
    assert not odd.warp_form
    with pytest.raises(cg.UnsupportedExpression):
        cg.generate_field(cg.FieldSpec(states, ["k"], [], odes, []), warp_form=True)


def test_constant_state_mask_and_adjoint_layout():
    """Host-side pieces of the save path / adjoint kernels: states with dy/dt == 0 are flagged
    (CTX_CONST_MASK), the shared-memory pack of the adjoint kernels fits 227 KB."""
    import sympy as sp

    from catalax_b200.neural import NeuralODE, RateFlowODE
    from catalax_b200.tools import codegen as cg, codegen_mlp as cm
    from tests import cases

    f = cg.generate_field(cases.field_spec(cases.michaelis_menten()))
    assert "#define CTX_CONST_MASK 0x1u" in f.source          # E (state 0 of [E, P, S]) is constant
    x, k = sp.symbols("x k")
    g = cg.generate_field(cg.FieldSpec(["x", "z"], ["k"], [], {"x": -k * x, "z": k * x - k * x}, []))
    assert "#define CTX_CONST_MASK 0x2u" in g.source          # z: identically zero after expansion
    for net in (NeuralODE(4, 64, 3, list("abcd"), activation="tanh", max_time=1.0, key=0),
                RateFlowODE(4, 3, 64, 3, list("abcd"), max_time=1.0, key=0)):
        d = cm.adjoint_dims(net.spec)
        assert d is not None and d["warps32"] >= 1 and d["warps64"] >= 1
        assert (d["spack"] + d["warps32"] * d["per_warp_bwd"]) * 4 <= 227 * 1024
        assert (d["spack"] + max(1, d["warps64"]) * d["per_warp_bwd"]) * 8 <= 227 * 1024
    big = NeuralODE(4, 256, 3, list("abcd"), max_time=1.0, key=0)   # 256-wide: does not fit
    src = cm.generate_mlp_field(big.spec).source
    assert "CTX_MLP_ADJ" not in src


def test_generated_mechanistic_vjp_matches_finite_differences():
    """UniversalODE adjoint (universalode.py:84-101): the generated adj_mech / adj_mech_vjp,
    compiled as host C++, against central differences."""
    import ctypes as C
    import hashlib
    import subprocess
    import tempfile
    from pathlib import Path

    import catalax_b200 as ctx
    from catalax_b200.neural import UniversalODE
    from catalax_b200.tools import codegen_mlp as cm

    model = ctx.Model(name="m")
    model.add_state(a="a", b="b")
    model.add_ode("a", "-k1 * a / (K + a) + 0.1 * exp(-t) * b")
    model.add_ode("b", "k1 * a / (K + a) - k2 * b * b + 0.01 * a_init")
    for n, v in (("k1", 0.8), ("k2", 0.3), ("K", 1.5)):
        model.parameters[n].value = v
    net = UniversalODE(2, 16, 1, model, max_time=1.0, key=0)
    lines = cm._adjoint_mech_source(net.spec, cm.layout_of(net.spec), cm.adjoint_dims(net.spec))
    body = "\n".join(l for l in lines if not l.startswith("#define"))
    src = ("#include <cmath>\ntypedef double real;\n#define R(x) x\n#define CTX_DIV(a, b) ((a) / (b))\n#define __device__\n"
           "#define __forceinline__ inline\nusing std::exp; using std::log; using std::sqrt;\n" + body +
           "\nextern \"C\" void f(double t, const double* y, const double* th, const double* y0, double* m)"
           " { adj_mech(t, y, th, y0, m); }\n"
           "extern \"C\" void v(double t, const double* y, const double* th, const double* y0,"
           " const double* fb, double* yb, double* tb) { adj_mech_vjp(t, y, th, y0, fb, yb, tb); }\n")
    d = Path(tempfile.gettempdir()) / "ctx_hostfield"
    d.mkdir(exist_ok=True)
    so = d / f"mech_{hashlib.sha1(src.encode()).hexdigest()[:12]}.so"
    if not so.exists():
        (d / (so.stem + ".cpp")).write_text(src)
        subprocess.run(["g++", "-O1", "-shared", "-fPIC", str(d / (so.stem + ".cpp")), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lib.f.argtypes = [C.c_double] + [C.POINTER(C.c_double)] * 4
    lib.v.argtypes = [C.c_double] + [C.POINTER(C.c_double)] * 6
    rng = np.random.default_rng(0)
    t, y, y0 = 0.7, rng.uniform(0.5, 2, 2), rng.uniform(0.5, 2, 2)
    th = np.asarray(net.parameters, dtype=np.float64).copy()
    fb = rng.normal(size=2)

    def mech(yy, tt):
        m = np.zeros(2)
        lib.f(t, P(np.ascontiguousarray(yy)), P(np.ascontiguousarray(tt)), P(y0), P(m))
        return m

    yb, tb = np.zeros(2), np.zeros(len(th))
    lib.v(t, P(y), P(th), P(y0), P(fb), P(yb), P(tb))
    for j in range(2):
        e = np.zeros(2); e[j] = 1e-6
        np.testing.assert_allclose(yb[j], fb @ (mech(y + e, th) - mech(y - e, th)) / 2e-6, rtol=1e-6, atol=1e-9)
    for p in range(len(th)):
        e = np.zeros(len(th)); e[p] = 1e-6
        np.testing.assert_allclose(tb[p], fb @ (mech(y, th + e) - mech(y, th - e)) / 2e-6, rtol=1e-6, atol=1e-9)
