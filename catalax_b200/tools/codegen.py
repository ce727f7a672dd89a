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
                s = f"CTX_DIV({s}, {' * '.join(den)})"
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
                    return prod if n > 0 else f"CTX_DIV({_lit(1.0)}, {prod})"
                return f"pow({self.p(b)}, {_lit(n)})"
            if x == sp.Rational(1, 2):
                return f"sqrt({self.p(b)})"
            if x == sp.Rational(-1, 2):
                return f"CTX_DIV({_lit(1.0)}, sqrt({self.p(b)}))"
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
    warp_form: bool = False   # lane-parallel form for the warp-per-trajectory kernel emitted
    warp_info: dict = field(default_factory=dict)


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


def _const_state_mask(f) -> int:
    """Bit i set <=> dy_i/dt == 0 identically (only reported for models of <= 32 states)."""
    if len(f) > 32:
        return 0
    mask = 0
    for i, e in enumerate(f):
        if e == 0 or sp.expand(e) == 0:
            mask |= 1 << i
    return mask


def generate_field(spec: FieldSpec, sens_theta: bool = False, sens_y0: bool = False,
                   warp_form="auto") -> GeneratedField:
    """Generate the CUDA device functions of one model.

    warp_form: "auto" emits the lane-parallel form (``generate_warp_form``) when a trajectory's
    state no longer fits the registers of one thread (S > 24) and the equations fall into few
    structural classes; True forces it (raises if impossible), False never emits it."""
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
        # states whose right-hand side is identically zero (conserved catalysts, ...): their
        # dense-output polynomial is the constant y, the save path skips the Horner evaluation
        f"#define CTX_CONST_MASK 0x{_const_state_mask(f):x}u",
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
    winfo = {}
    if warp_form is True or (warp_form == "auto" and S > WARP_FORM_MIN_STATES):
        wsrc, winfo = generate_warp_form(spec)
        if wsrc is None:
            if warp_form is True:
                raise UnsupportedExpression(f"no lane-parallel form: {winfo.get('reason')}")
            winfo = {}
        else:
            src += wsrc
    text = "\n".join(src) + "\n"
    key = hashlib.sha256(text.encode()).hexdigest()[:24]
    return GeneratedField(text, S, P, C, D, n_theta, D - n_theta, flops_rhs, flops_tan, uses_t, key,
                          bool(winfo), winfo)


# ---------------------------------------------------------------- lane-parallel ("warp") form
WARP_FORM_MIN_STATES = 24     # below this a trajectory lives in the registers of one thread
WARP_FORM_MAX_TERM_SLOTS = 12


def _structural_key(expr: sp.Expr, kinds: Dict[sp.Symbol, Tuple[str, int]]):
    """Rename the leaves of ``expr`` to typed placeholders in order of first appearance.

    Returns (key, template expression over placeholder symbols, [(kind, index)] per placeholder).
    Two equations with the same key differ only in WHICH state / parameter / constant / initial
    value they read, so the 32 lanes of a warp can evaluate 32 of them with one instruction
    stream and per-lane index tables."""
    order: List[sp.Symbol] = []
    for node in sp.preorder_traversal(expr):
        if isinstance(node, sp.Symbol) and node != T_SYM and node not in order:
            if node not in kinds:
                raise KeyError(f"Missing input for symbol {node}")
            order.append(node)
    counters: Dict[str, int] = {}
    sub, leaves = {}, []
    for sym in order:
        kind, index = kinds[sym]
        n = counters.get(kind, 0)
        counters[kind] = n + 1
        sub[sym] = sp.Symbol(f"__{kind}{n}")
        leaves.append((kind, index))
    templ = expr.xreplace(sub)
    return sp.srepr(templ), templ, [sub[s] for s in order], leaves


def generate_warp_form(spec: FieldSpec):
    """Lane-parallel form of the vector field for the warp-per-trajectory kernel ``ctx_solve_w``.

    The field is the same ``dy = odes + S_mat @ rates`` as ``ctx_rhs`` (stack.py:159-173,
    232-250, 332-349) but laid out for 32 lanes working on ONE trajectory:

    * every non-zero ODE right-hand side and every reaction rate is a *term*; terms are grouped
      by structural class (``_structural_key``) and each class is padded to whole slots of 32
      terms, so slot q of lane l evaluates term ``32 q + l`` with straight-line code shared by
      all lanes (state values gathered from shared memory through per-lane index registers;
      parameters / constants / initial values pre-loaded into per-lane registers once per
      trajectory);
    * ``dy_i = sum_k coef[i][k] * term[col[i][k]]`` in ELL format (rows padded to the slot's
      longest row with a column that always holds 0), lane l owning states ``l, l+32, ...``.

    Returns (lines, info) or (None, {"reason": ...}) when the equations do not share structure.
    """
    S = len(spec.states)
    kinds: Dict[sp.Symbol, Tuple[str, int]] = {}
    for i, s_ in enumerate(spec.states):
        kinds[sp.Symbol(s_)] = ("Y", i)
        kinds[sp.Symbol(f"{s_}_init")] = ("I", i)
    for i, p_ in enumerate(spec.parameters):
        kinds[sp.Symbol(p_)] = ("P", i)
    for i, c_ in enumerate(spec.constants):
        kinds[sp.Symbol(c_)] = ("C", i)
    sidx = {s_: i for i, s_ in enumerate(spec.states)}

    # ---- terms: (expression, {state index: coefficient})
    terms = []
    for s_, e in spec.odes.items():
        e = sp.sympify(e)
        if e != 0 and s_ in sidx:
            terms.append((e, {sidx[s_]: 1.0}))
    for _sym, reactants, products, expr in sorted(spec.reactions, key=lambda r: r[0]):
        row: Dict[int, float] = {}
        for coef, st in reactants:
            if st in sidx:
                row[sidx[st]] = row.get(sidx[st], 0.0) - float(np.float32(coef))
        for coef, st in products:
            if st in sidx:
                row[sidx[st]] = row.get(sidx[st], 0.0) + float(np.float32(coef))
        row = {i: v for i, v in row.items() if v != 0.0}
        if row:
            terms.append((sp.sympify(expr), row))

    # ---- structural classes -> slots
    classes: Dict[str, dict] = {}
    for n, (e, _row) in enumerate(terms):
        key, templ, placeholders, leaves = _structural_key(e, kinds)
        cl = classes.setdefault(key, {"templ": templ, "ph": placeholders, "members": []})
        cl["members"].append((n, leaves))
    n_term_slots = sum((len(c["members"]) + 31) // 32 for c in classes.values())
    if n_term_slots > WARP_FORM_MAX_TERM_SLOTS:
        return None, {"reason": f"{len(classes)} structural classes need {n_term_slots} term slots"}
    NSS = (S + 31) // 32
    zero_col = n_term_slots * 32          # rs[zero_col] == 0 always

    itab: List[List[int]] = []            # [entry][lane]
    rtab: List[List[float]] = []
    pre_lines, term_lines = [], []
    pos_of_term: Dict[int, int] = {}
    n_pre = 0
    flops = 0
    slot = 0
    for cl in classes.values():
        members = cl["members"]
        for s0 in range(0, len(members), 32):
            chunk = members[s0:s0 + 32]
            names: Dict[sp.Symbol, str] = {T_SYM: "t"}
            for q, ph in enumerate(cl["ph"]):
                col = [chunk[l][1][q][1] if l < len(chunk) else 0 for l in range(32)]
                kind = chunk[0][1][q][0]
                e_idx = len(itab)
                itab.append(col)
                if kind == "Y":
                    names[ph] = f"ys[it[{e_idx}]]"
                else:
                    src_arr = {"P": "th", "C": "c", "I": "y0"}[kind]
                    pre_lines.append(f"  pre[{n_pre}] = {src_arr}[it[{e_idx}]];")
                    names[ph] = f"pre[{n_pre}]"
                    n_pre += 1
            lines, fl = _emit_block([(f"rs[{slot * 32} + lane]", cl["templ"])], names, f"w{slot}_")
            term_lines += lines
            flops += fl * len(chunk)
            for l, (n, _leaves) in enumerate(chunk):
                pos_of_term[n] = slot * 32 + l
            slot += 1

    # ---- ELL rows of dy = S_mat @ terms
    rows: List[List[Tuple[int, float]]] = [[] for _ in range(NSS * 32)]
    for n, (_e, row) in enumerate(terms):
        for i, v in row.items():
            rows[i].append((pos_of_term[n], v))
    for r in rows:                          # equal coefficients line up across lanes -> literals
        r.sort(key=lambda cv: (cv[1], cv[0]))
    dy_lines = []
    for u in range(NSS):
        width = max((len(rows[u * 32 + l]) for l in range(32)), default=0)
        parts = []
        for k in range(width):
            cols, coefs = [], []
            for l in range(32):
                r = rows[u * 32 + l]
                cols.append(r[k][0] if k < len(r) else zero_col)
                coefs.append(r[k][1] if k < len(r) else 0.0)
            ei = len(itab)
            itab.append(cols)
            live = {v for v, cc in zip(coefs, cols) if cc != zero_col}
            if len(live) == 1:              # same coefficient in every lane that has the entry
                parts.append(f"{_lit(live.pop())} * rs[it[{ei}]]")
            else:
                parts.append(f"rt[{len(rtab)}] * rs[it[{ei}]]")
                rtab.append(coefs)
        dy_lines.append(f"  dy[{u}] = " + (" + ".join(parts) if parts else "R(0.0)") + ";")
        flops += 2 * sum(len(rows[u * 32 + l]) for l in range(32))

    NI, NR = max(len(itab), 1), max(len(rtab), 1)
    if not itab:
        itab = [[0] * 32]
    if not rtab:
        rtab = [[0.0] * 32]
    per_warp_bytes = (NSS * 32 + n_term_slots * 32 + 32) * 8
    warps = max(1, min(8, (40 * 1024) // per_warp_bytes))
    if per_warp_bytes > 40 * 1024:
        return None, {"reason": "term table does not fit shared memory"}
    fl = lambda v: _lit(v)
    out = [
        "#define CTX_HAVE_WARP 1",
        f"#define CTXW_NSS {NSS}",
        f"#define CTXW_NTS {n_term_slots}",
        f"#define CTXW_NI {NI}",
        f"#define CTXW_NR {NR}",
        f"#define CTXW_NPRE {max(n_pre, 1)}",
        f"#define CTXW_WARPS {warps}",
        "__device__ const int ctxw_itab[CTXW_NI * 32] = {",
        *["  " + ", ".join(str(v) for v in row) + "," for row in itab],
        "};",
        "__device__ const real ctxw_rtab[CTXW_NR * 32] = {",
        *["  " + ", ".join(fl(v) for v in row) + "," for row in rtab],
        "};",
        "__device__ __forceinline__ void ctxw_preload(const int* it, const real* th, const real* c,",
        "                                             const real* y0, real* pre) {",
        "  (void)it; (void)th; (void)c; (void)y0; (void)pre;",
        *pre_lines,
        "}",
        "__device__ __forceinline__ void ctxw_terms(const int* it, const real* pre, const real t,",
        "                                           const real* ys, real* rs, const int lane) {",
        "  (void)it; (void)pre; (void)t; (void)ys;",
        *term_lines,
        "}",
        "__device__ __forceinline__ void ctxw_dy(const int* it, const real* rt, const real* rs,",
        "                                        real* dy) {",
        "  (void)it; (void)rt; (void)rs;",
        *dy_lines,
        "}",
    ]
    info = {"classes": len(classes), "term_slots": n_term_slots, "state_slots": NSS,
            "terms": len(terms), "flops_rhs": flops, "warps": warps,
            "lane_utilisation": S / (NSS * 32.0)}
    return out, info
