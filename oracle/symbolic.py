"""CPU restatement of the reference's right-hand-side stacks.  PARITY ORACLE (test infra).

Follows ``catalax/tools/stack.py``:

* ``ODEStack.__call__``       (:159-173)  -> one expression per state, stacked;
* ``ReactionStack.__call__``  (:232-250)  -> ``stoich_mat @ rates``;
* ``MixedStack.__call__``     (:332-349)  -> sum of both, rows without an equation are 0;
* symbol resolution of ``catalax/tools/symbolicmodule.py:136-159,266-290``: states,
  parameters, constants and ``{state}_init`` symbols are looked up by POSITION in the
  arrays ``y``, ``theta``, ``constants``, ``y0``; ``t`` by name (``stack.py:94-96``);
* stoichiometric coefficients are cast to float32 (``catalax/model/stoich.py:195``),
  reactions are ordered by symbol (``stoich.py:47-59``).

It deliberately does NOT share code with ``catalax_b200/tools/codegen.py``: expressions
are evaluated through ``sympy.lambdify`` (numpy), with no CSE and no hand-written printer.
"""
from __future__ import annotations

import numpy as np
import sympy as sp

T_SYM = sp.Symbol("t")


def _syms(names):
    return [sp.Symbol(n) for n in names]


def total_rhs_exprs(states, odes, reactions):
    """Per-state sympy expression of dy/dt (ODE + stoichiometry * rate).

    states: ordered list of modeled state names; odes: {state: expr};
    reactions: list of (symbol, [(coef, state)...] reactants, [(coef, state)...] products, expr).
    """
    total = {s: sp.Integer(0) for s in states}
    for s, e in odes.items():
        total[s] = total[s] + sp.sympify(e)
    for _sym, reactants, products, expr in sorted(reactions, key=lambda r: r[0]):
        expr = sp.sympify(expr)
        for coef, st in reactants:
            if st in total:
                total[st] = total[st] - sp.Float(float(np.float32(coef))) * expr
        for coef, st in products:
            if st in total:
                total[st] = total[st] + sp.Float(float(np.float32(coef))) * expr
    return [total[s] for s in states]


def make_rhs(states, parameters, constants, odes, reactions=(), sens_theta=False,
             sens_y0=False):
    """Return (rhs, N, D) with rhs(t[B], Y[B,N], theta[B,P], consts[B,C], y0[B,S]) -> [B,N].

    With sensitivities the state is augmented direction-major:
    ``Y = [y, dy/dtheta_0, ..., dy/dtheta_{P-1}, dy/dy0_0, ..., dy/dy0_{S-1}]`` (each S long).
    """
    S, P = len(states), len(parameters)
    y_s = _syms(states)
    th_s = _syms(parameters)
    c_s = _syms(constants)
    i_s = _syms([f"{s}_init" for s in states])
    exprs = total_rhs_exprs(states, dict(odes), list(reactions))
    args = [T_SYM, *y_s, *th_s, *c_s, *i_s]
    f_fun = [sp.lambdify(args, e, modules="numpy") for e in exprs]

    dirs = []
    if sens_theta:
        dirs += [("theta", p) for p in range(P)]
    if sens_y0:
        dirs += [("y0", i) for i in range(S)]
    D = len(dirs)
    jac = [[sp.diff(e, y) for y in y_s] for e in exprs]
    jac_fun = [[None if j == 0 else sp.lambdify(args, j, modules="numpy") for j in row]
               for row in jac]
    force = []
    for kind, idx in dirs:
        sym = th_s[idx] if kind == "theta" else i_s[idx]
        row = [sp.diff(e, sym) for e in exprs]
        force.append([None if g == 0 else sp.lambdify(args, g, modules="numpy") for g in row])

    def rhs(t, Y, theta, consts, y0):
        B = Y.shape[0]
        vals = [t] + [Y[:, i] for i in range(S)] + [theta[:, p] for p in range(P)] \
            + [consts[:, c] for c in range(len(constants))] + [y0[:, i] for i in range(S)]
        out = np.empty_like(Y)
        for i in range(S):
            out[:, i] = np.broadcast_to(f_fun[i](*vals), (B,))
        if D:
            J = [[None if jf is None else np.broadcast_to(jf(*vals), (B,)) for jf in row]
                 for row in jac_fun]
            for d in range(D):
                s_d = Y[:, S + d * S: S + (d + 1) * S]
                for i in range(S):
                    acc = np.zeros(B, dtype=Y.dtype)
                    for j in range(S):
                        if J[i][j] is not None:
                            acc = acc + J[i][j] * s_d[:, j]
                    if force[d][i] is not None:
                        acc = acc + np.broadcast_to(force[d][i](*vals), (B,))
                    out[:, S + d * S + i] = acc
        return out

    def init_aug(B, dtype):
        aug = np.zeros((B, D * S), dtype=dtype)
        for d, (kind, idx) in enumerate(dirs):
            if kind == "y0":
                aug[:, d * S + idx] = 1
        return aug

    return rhs, S * (1 + D), D, init_aug
