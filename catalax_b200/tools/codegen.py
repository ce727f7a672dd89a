"""sympy -> CUDA device-function code generation for a kinetic model's vector field.

Replaces the reference's expression-tree interpreter ``catalax/tools/symbolicmodule.py``
(op table :52-98, positional symbol lookup :136-159, node builder :266-290) and the three
RHS stacks of ``catalax/tools/stack.py`` (ODEStack :159-173, ReactionStack :232-250,
MixedStack :332-349): instead of tracing an expression tree through JAX we print one CUDA
``__device__`` function per model, common sub-expressions eliminated, which NVRTC compiles
for sm_100a together with the integrator template (``csrc/ctx_integrator.cuh``).

Symbol resolution follows the reference exactly: states / parameters / constants /
``{state}_init`` are positional (index into ``y``, ``th``, ``c``, ``y0``); ``t`` is by name.
The reaction form evaluates each rate once and applies the stoichiometric matrix
(coefficients rounded to float32 first, ``catalax/model/stoich.py:195``) as explicit
sparse sums; rows without an equation are zero (``stack.py:175-200,252-281``).

Besides ``ctx_rhs`` the generator emits ``ctx_rhs_tan``: the vector field together with the
tangent field ``J(y) s_d + df/d(dir_d)`` for D sensitivity directions (parameters and/or
initial values).  Carried through the same Runge-Kutta stages with the step sizes frozen this
is exactly what ``jax.jacobian`` / ``value_and_grad`` of the reference's solve computes
(``simulation.py:279-323``; numpyro's gradient of ``mcmc.py:448``).
"""
from __future__ import annotations

import hashlib
import re
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np
import sympy as sp
from sympy.printing.precedence import precedence

T_SYM = sp.Symbol("t")


class UnsupportedExpression(ValueError):
    pass


# ---------------------------------------------------------------- expression printer
_FUNCS = {
    sp.exp: "exp", sp.log: "log", sp.sin: "sin", sp.cos: "cos", sp.tan: "tan",
    sp.asin: "asin", sp.acos: "acos", sp.atan: "atan", sp.sinh: "sinh", sp.cosh: "cosh",
    sp.tanh: "tanh", sp.asinh: "asinh", sp.acosh: "acosh", sp.atanh: "atanh",
    sp.erf: "erf", sp.Abs: "fabs", sp.floor: "floor", sp.ceiling: "ceil",
    sp.atan2: "atan2",
}
_RELS = {
    sp.Eq: "==", sp.Ne: "!=", sp.StrictGreaterThan: ">", sp.StrictLessThan: "<",
    sp.GreaterThan: ">=", sp.LessThan: "<=",
}


def _lit(x: float) -> str:
    r = repr(float(x))
    if r in ("inf", "-inf", "nan"):
        raise UnsupportedExpression(f"non-finite literal {r}")
    if "." not in r and "e" not in r and "E" not in r:
        r += ".0"
    return f"R({r})"


_ATOM = re.compile(r"^(R\([^()]*\)|[A-Za-z_][A-Za-z0-9_]*(\[[0-9]+\])?)$")


def _is_atom(s: str) -> bool:
    return bool(_ATOM.match(s))


class _Printer:
    """Prints a sympy expression as a C++ expression over ``real``."""

    def __init__(self, names: Dict[sp.Symbol, str]):
        self.names = names

    def p(self, e) -> str:
        if isinstance(e, sp.Symbol):
            try:
                return self.names[e]
            except KeyError as err:
                # same failure mode as symbolicmodule.py:125-129
                raise KeyError(f"Missing input for symbol {e}") from err
        if isinstance(e, sp.Integer):
            return _lit(int(e))
        if isinstance(e, (sp.Float, sp.Rational)):
            return _lit(float(e))
        if e is sp.pi or e is sp.E or isinstance(e, sp.NumberSymbol):
            return _lit(float(e))
        if isinstance(e, sp.Add):
            out = ""
            for i, a in enumerate(e.as_ordered_terms()):
                c, rest = a.as_coeff_Mul()
                if i and c < 0:
                    out += " - " + self._paren(-a)
                elif i:
                    out += " + " + self._paren(a)
                else:
                    out += self._paren(a)
            return out
        if isinstance(e, sp.Mul):
            coeff, rest = e.as_coeff_Mul()
            num, den = [], []
            if coeff == -1:
                sign = "-"
            else:
                sign = ""
                if coeff != 1:
                    num.append(_lit(float(coeff)))
            for f in sp.Mul.make_args(rest):
                if isinstance(f, sp.Pow) and f.exp.is_number and f.exp < 0:
                    den.append(self._paren(sp.Pow(f.base, -f.exp), mul=True))
                else:
                    num.append(self._paren(f, mul=True))
            s = " * ".join(num) if num else _lit(1.0)
            if den:
                d = " * ".join(den)
                s = f"{s} / {d}" if _is_atom(d) else f"{s} / ({d})"
            return f"{sign}({s})" if sign else s
        if isinstance(e, sp.Pow):
            b, x = e.base, e.exp
            if x.is_Integer:
                n = int(x)
                if n == 0:
                    return _lit(1.0)
                bs = self.p(b)
                if not _is_atom(bs):
                    bs = f"({bs})"
                if 1 <= abs(n) <= 4:
                    prod = " * ".join([bs] * abs(n))
                    return prod if n > 0 else f"{_lit(1.0)} / ({prod})"
                return f"pow({self.p(b)}, {_lit(n)})"
            if x == sp.Rational(1, 2):
                return f"sqrt({self.p(b)})"
            if x == sp.Rational(-1, 2):
                return f"{_lit(1.0)} / sqrt({self.p(b)})"
            return f"pow({self.p(b)}, {self.p(x)})"
        if isinstance(e, sp.sign):
            a = self.p(e.args[0])
            return f"ctx_sign({a})"
        if isinstance(e, (sp.Max, sp.Min)):
            fn = "fmax" if isinstance(e, sp.Max) else "fmin"
            args = [self.p(a) for a in e.args]
            out = args[0]
            for a in args[1:]:
                out = f"{fn}({out}, {a})"
            return out
        if type(e) in _RELS:
            return f"(({self.p(e.args[0])}) {_RELS[type(e)]} ({self.p(e.args[1])}) ? R(1.0) : R(0.0))"
        if isinstance(e, (sp.And, sp.Or, sp.Xor)):
            op = {"And": "&&", "Or": "||", "Xor": "!="}[type(e).__name__]
            parts = [f"(({self.p(a)}) != R(0.0))" for a in e.args]
            return "((" + f" {op} ".join(parts) + ") ? R(1.0) : R(0.0))"
        if isinstance(e, sp.Not):
            return f"((({self.p(e.args[0])}) != R(0.0)) ? R(0.0) : R(1.0))"
        if isinstance(e, (sp.re,)):
            return self.p(e.args[0])
        for cls, name in _FUNCS.items():
            if isinstance(e, cls):
                return f"{name}(" + ", ".join(self.p(a) for a in e.args) + ")"
        raise UnsupportedExpression(f"Unsupported Sympy type {type(e)}")

    def _paren(self, e, mul=False) -> str:
        s = self.p(e)
        if isinstance(e, sp.Add) or (mul and isinstance(e, sp.Mul) and s.startswith("-")):
            return f"({s})"
        if not mul and isinstance(e, sp.Mul):
            return s
        return s


# ---------------------------------------------------------------- model description
@dataclass
class FieldSpec:
    """Canonical description of a mechanistic vector field (all names already ordered)."""

    states: List[str]
    parameters: List[str]
    constants: List[str]
    odes: Dict[str, sp.Expr] = field(default_factory=dict)
    # (symbol, [(coef, state)], [(coef, state)], rate expression), any order
    reactions: List[Tuple[str, list, list, sp.Expr]] = field(default_factory=list)

    def symbol_names(self) -> Dict[sp.Symbol, str]:
        names = {T_SYM: "t"}
        for i, s in enumerate(self.states):
            names[sp.Symbol(s)] = f"y[{i}]"
            names[sp.Symbol(f"{s}_init")] = f"y0[{i}]"
        for i, p in enumerate(self.parameters):
            names[sp.Symbol(p)] = f"th[{i}]"
        for i, c in enumerate(self.constants):
            names[sp.Symbol(c)] = f"c[{i}]"
        return names

    def total_exprs(self) -> List[sp.Expr]:
        """dy_i/dt = ode_i + sum_j S_ij * rate_j with f32-rounded stoichiometry."""
        total = {s: sp.Integer(0) for s in self.states}
        for s, e in self.odes.items():
            total[s] = total[s] + e
        for _sym, reactants, products, expr in sorted(self.reactions, key=lambda r: r[0]):
            for coef, st in reactants:
                if st in total:
                    total[st] = total[st] - _coef(coef) * expr
            for coef, st in products:
                if st in total:
                    total[st] = total[st] + _coef(coef) * expr
        return [total[s] for s in self.states]


def _coef(c) -> sp.Expr:
    v = float(np.float32(c))
    return sp.Integer(int(v)) if v == int(v) else sp.Float(v)


@dataclass
class GeneratedField:
    source: str          # CUDA text defining ctx_rhs (and ctx_rhs_tan when D > 0)
    S: int
    P: int
    C: int
    D: int               # number of sensitivity directions
    n_theta_dirs: int
    n_y0_dirs: int
    flops_rhs: int       # sympy.count_ops of the CSE'd field (roofline accounting)
    flops_tan: int
    uses_t: bool
    key: str             # content hash (NVRTC cache key)


def _emit_block(exprs_out: Sequence[Tuple[str, sp.Expr]], names, tmp_prefix: str):
    """CSE a list of (lhs, expr) and return (lines, flop count)."""
    targets = [e for _, e in exprs_out]
    syms = sp.numbered_symbols(tmp_prefix)
    repl, reduced = sp.cse(targets, symbols=syms, optimizations="basic", order="none")
    local = dict(names)
    pr = _Printer(local)
    lines = []
    flops = 0
    for s, e in repl:
        local[s] = str(s)
        lines.append(f"  const real {s} = {pr.p(e)};")
        flops += int(sp.count_ops(e))
    for (lhs, _), e in zip(exprs_out, reduced):
        lines.append(f"  {lhs} = {pr.p(e)};")
        flops += int(sp.count_ops(e))
    return lines, flops


def generate_field(spec: FieldSpec, sens_theta: bool = False, sens_y0: bool = False) -> GeneratedField:
    """Generate the CUDA device functions of one model."""
    S, P, C = len(spec.states), len(spec.parameters), len(spec.constants)
    names = spec.symbol_names()
    f = [sp.sympify(e) for e in spec.total_exprs()]
    for e in f:
        bad = [s for s in e.free_symbols if s not in names]
        if bad:
            raise KeyError(f"Missing input for symbol {bad[0]}")
    uses_t = any(T_SYM in e.free_symbols for e in f)

    lines, flops_rhs = _emit_block([(f"dy[{i}]", e) for i, e in enumerate(f)], names, "x")
    src = [
        f"#define CTX_S {S}",
        f"#define CTX_P {P}",
        f"#define CTX_C {C}",
        "__device__ __forceinline__ void ctx_rhs(const real t, const real* y, const real* th,",
        "                                        const real* c, const real* y0, real* dy) {",
        "  (void)t; (void)th; (void)c; (void)y0;",
        *lines,
        "}",
    ]

    dirs = []
    if sens_theta:
        dirs += [sp.Symbol(p) for p in spec.parameters]
    n_theta = len(dirs)
    if sens_y0:
        dirs += [sp.Symbol(f"{s}_init") for s in spec.states]
    D = len(dirs)
    flops_tan = 0
    if D:
        y_syms = [sp.Symbol(s) for s in spec.states]
        outs = [(f"dy[{i}]", e) for i, e in enumerate(f)]
        jac_nz = []
        for i, e in enumerate(f):
            for j, ys in enumerate(y_syms):
                d = sp.diff(e, ys)
                if d != 0:
                    jac_nz.append((i, j))
                    outs.append((f"const real J{i}_{j}", d))
        frc_nz = []
        for d_i, sym in enumerate(dirs):
            for i, e in enumerate(f):
                g = sp.diff(e, sym)
                if g != 0:
                    frc_nz.append((d_i, i))
                    outs.append((f"const real G{d_i}_{i}", g))
        tl, flops_tan = _emit_block(outs, names, "z")
        body = list(tl)
        body.append("  #pragma unroll")
        body.append("  for (int d = 0; d < CTX_D; ++d) {")
        for i in range(S):
            terms = [f"J{i}_{j} * s[d * CTX_S + {j}]" for (ii, j) in jac_nz if ii == i]
            body.append(f"    ds[d * CTX_S + {i}] = " + (" + ".join(terms) if terms else "R(0.0)") + ";")
            flops_tan += 2 * len(terms) * D
        body.append("  }")
        for d_i, i in frc_nz:
            body.append(f"  ds[{d_i} * CTX_S + {i}] += G{d_i}_{i};")
            flops_tan += 1
        src += [
            f"#define CTX_D {D}",
            f"#define CTX_D_THETA {n_theta}",
            f"#define CTX_D_Y0 {D - n_theta}",
            "__device__ __forceinline__ void ctx_rhs_tan(const real t, const real* y, const real* s,",
            "                                            const real* th, const real* c, const real* y0,",
            "                                            real* dy, real* ds) {",
            "  (void)t; (void)th; (void)c; (void)y0;",
            *body,
            "}",
        ]
    else:
        src += ["#define CTX_D 0", "#define CTX_D_THETA 0", "#define CTX_D_Y0 0"]
    text = "\n".join(src) + "\n"
    key = hashlib.sha256(text.encode()).hexdigest()[:24]
    return GeneratedField(text, S, P, C, D, n_theta, D - n_theta, flops_rhs, flops_tan, uses_t, key)
