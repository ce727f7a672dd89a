"""Compiled (C + OpenMP) form of the oracle: the CPU baseline and the large-size checker.

TEST INFRASTRUCTURE ONLY.  Same algorithm and operation order as ``diffrax_oracle.py`` (Tsit5
path), one trajectory per OpenMP task; the model's right-hand side is printed by sympy's OWN C
printer (``sympy.ccode``) -- no code is shared with the product's CUDA generator
(``catalax_b200/tools/codegen.py``).

kind = "port" in bench.py's ``cpu_baseline``: this is a restatement of the reference's
algorithm (diffrax cannot be installed here), timed on the GPU box's host cores.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np
import sympy as sp

from .symbolic import T_SYM, total_rhs_exprs

HERE = Path(__file__).resolve().parent
BUILD = HERE / "_build"


def _augmented_exprs(states, parameters, constants, odes, reactions, sens_theta, sens_y0):
    S = len(states)
    y_s = [sp.Symbol(s) for s in states]
    th_s = [sp.Symbol(p) for p in parameters]
    i_s = [sp.Symbol(f"{s}_init") for s in states]
    f = total_rhs_exprs(states, dict(odes), list(reactions))
    dirs = ([("theta", p) for p in range(len(parameters))] if sens_theta else []) + \
           ([("y0", i) for i in range(S)] if sens_y0 else [])
    exprs = list(f)
    tan_syms = {}
    for d, (kind, idx) in enumerate(dirs):
        sym = th_s[idx] if kind == "theta" else i_s[idx]
        for i in range(S):
            e = sp.diff(f[i], sym)
            for j in range(S):
                tj = tan_syms.setdefault((d, j), sp.Symbol(f"ctxtan_{d}_{j}"))
                e = e + sp.diff(f[i], y_s[j]) * tj
            exprs.append(e)
    return exprs, dirs, tan_syms


def _source(states, parameters, constants, odes, reactions, sens_theta, sens_y0, f32):
    S = len(states)
    exprs, dirs, tan_syms = _augmented_exprs(states, parameters, constants, odes, reactions,
                                             sens_theta, sens_y0)
    N = len(exprs)
    sub = {T_SYM: sp.Symbol("t")}
    names = {}
    for i, s in enumerate(states):
        names[sp.Symbol(s)] = sp.Symbol(f"Y[{i}]")
        names[sp.Symbol(f"{s}_init")] = sp.Symbol(f"y0[{i}]")
    for i, p in enumerate(parameters):
        names[sp.Symbol(p)] = sp.Symbol(f"th[{i}]")
    for i, c in enumerate(constants):
        names[sp.Symbol(c)] = sp.Symbol(f"c[{i}]")
    for (d, j), sym in tan_syms.items():
        names[sym] = sp.Symbol(f"Y[{S + d * S + j}]")
    lines = []
    for i, e in enumerate(exprs):
        code = sp.ccode(sp.sympify(e).xreplace(names))
        lines.append(f"  dY[{i}] = {code};")
    real = "float" if f32 else "double"
    body = "\n".join(lines)
    return f"""
#define ORACLE_F32 {1 if f32 else 0}
typedef {real} real;
#define ORACLE_S {S}
#define ORACLE_N {N}
#define ORACLE_NERR {S}
#include <math.h>
static void oracle_rhs(real t, const real* Y, const real* th, const real* c, const real* y0,
                       real* dY) {{
  (void)t; (void)th; (void)c; (void)y0;
{body}
}}
#include "ctx_oracle_core.h"
""", N, dirs


class COracle:
    """A model compiled into oracle/_build/*.so."""

    def __init__(self, states, parameters, constants, odes, reactions=(), sens_theta=False,
                 sens_y0=False, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        f32 = self.dtype == np.float32
        src, self.N, self.dirs = _source(list(states), list(parameters), list(constants),
                                         dict(odes), list(reactions), sens_theta, sens_y0, f32)
        self.S, self.P, self.C = len(states), len(parameters), len(constants)
        core = (HERE / "ctx_oracle_core.h").read_text()
        key = hashlib.sha256((src + core).encode()).hexdigest()[:20]
        BUILD.mkdir(exist_ok=True)
        so = BUILD / f"oracle_{key}.so"
        if not so.exists():
            cfile = BUILD / f"oracle_{key}.c"
            cfile.write_text(src)
            tmp = BUILD / f"oracle_{key}.{os.getpid()}.tmp.so"
            # -ffp-contract=off: no FMA contraction, so f64 results equal the numpy oracle's
            subprocess.run(["gcc", "-O2", "-march=native", "-ffp-contract=off", "-fopenmp",
                            "-shared", "-fPIC", f"-I{HERE}", str(cfile), "-o", str(tmp), "-lm"],
                           check=True)
            os.replace(tmp, so)
        self.lib = C.CDLL(str(so))
        vp = C.c_void_p
        self.lib.oracle_solve.argtypes = [C.c_int64, C.c_int, vp, C.c_int64, vp, C.c_int64, vp,
                                          C.c_int64, vp, C.c_int64, vp, C.c_double, C.c_double,
                                          C.c_double, C.c_int, vp, vp, vp, vp, C.c_int]
        self.lib.oracle_solve.restype = None

    def init_aug(self):
        aug = np.zeros(self.N - self.S, dtype=self.dtype)
        for d, (kind, idx) in enumerate(self.dirs):
            if kind == "y0":
                aug[d * self.S + idx] = 1
        return aug

    def solve(self, y0, theta, consts, ts, *, rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096,
              n_threads=None):
        dt = self.dtype
        y0 = np.ascontiguousarray(y0, dtype=dt); theta = np.ascontiguousarray(theta, dtype=dt)
        consts = np.ascontiguousarray(consts, dtype=dt); ts = np.ascontiguousarray(ts, dtype=dt)
        B = max([a.shape[0] for a in (y0, theta, consts, ts) if a.ndim == 2] or [1])
        T = ts.shape[-1]
        st = lambda a, w: w if a.ndim == 2 else 0
        ys = np.empty((B, T, self.N), dtype=dt)
        status = np.empty(B, dtype=np.int32); nacc = np.empty(B, dtype=np.int32)
        nrej = np.empty(B, dtype=np.int32)
        aug = self.init_aug()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        n_threads = n_threads or os.cpu_count() or 1
        self.lib.oracle_solve(B, T, p(y0), st(y0, self.S), p(theta), st(theta, self.P), p(consts),
                              st(consts, self.C), p(ts), st(ts, T),
                              p(aug) if aug.size else None, float(rtol), float(atol), float(dt0),
                              int(max_steps), p(ys), p(status), p(nacc), p(nrej), int(n_threads))
        from .diffrax_oracle import OracleResult

        return OracleResult(ys, status, nacc, nrej)


def build():
    """Pre-compile the oracle libraries of the benchmark models (called by build())."""
    import sys

    root = HERE.parent
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    from tests import cases

    m = cases.michaelis_menten()
    for dtype in (np.float32, np.float64):
        COracle(*m[:4], m[4], dtype=dtype)
        COracle(*m[:4], m[4], sens_theta=True, dtype=dtype)
