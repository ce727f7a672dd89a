"""Compile a code-generated vector field as HOST C++ (g++) so the generator can be checked on
a machine without a GPU.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
import tempfile
from pathlib import Path

import numpy as np

_TEMPLATE = r"""
#include <cmath>
typedef double real;
#define R(x) x
#define CTX_DIV(a, b) ((a) / (b))
#define __device__
#define __forceinline__ inline
using std::fmax; using std::fmin; using std::fabs;
static inline real ctx_sign(real x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }
%s
extern "C" void eval_rhs(double t, const double* y, const double* th, const double* c,
                         const double* y0, double* dy) { ctx_rhs(t, y, th, c, y0, dy); }
#if CTX_D > 0
extern "C" void eval_tan(double t, const double* y, const double* s, const double* th,
                         const double* c, const double* y0, double* dy, double* ds) {
  ctx_rhs_tan(t, y, s, th, c, y0, dy, ds);
}
#endif
#ifdef CTX_HAVE_WARP
// the lane-parallel form, emulated lane by lane exactly as ctx_solve_w runs it
extern "C" void eval_warp(double t, const double* y, const double* th, const double* c,
                          const double* y0, double* dy) {
  static double ys[CTXW_NSS * 32], rs[CTXW_NTS * 32 + 32];
  for (int i = 0; i < CTXW_NSS * 32; ++i) ys[i] = i < CTX_S ? y[i] : 0.0;
  for (int i = 0; i < CTXW_NTS * 32 + 32; ++i) rs[i] = 0.0;
  int it[CTXW_NI]; double rt[CTXW_NR], pre[CTXW_NPRE];
  for (int lane = 0; lane < 32; ++lane) {
    for (int k = 0; k < CTXW_NI; ++k) it[k] = ctxw_itab[k * 32 + lane];
    ctxw_preload(it, th, c, y0, pre);
    ctxw_terms(it, pre, t, ys, rs, lane);
  }
  for (int i = CTXW_NTS * 32; i < CTXW_NTS * 32 + 32; ++i) rs[i] = 0.0;
  for (int lane = 0; lane < 32; ++lane) {
    for (int k = 0; k < CTXW_NI; ++k) it[k] = ctxw_itab[k * 32 + lane];
    for (int k = 0; k < CTXW_NR; ++k) rt[k] = ctxw_rtab[k * 32 + lane];
    double d[CTXW_NSS];
    ctxw_dy(it, rt, rs, d);
    for (int u = 0; u < CTXW_NSS; ++u)
      if (u * 32 + lane < CTX_S) dy[u * 32 + lane] = d[u];
  }
}
#endif
"""


class HostField:
    def __init__(self, field):
        src = _TEMPLATE % field.source
        key = hashlib.sha1(src.encode()).hexdigest()[:16]
        d = Path(tempfile.gettempdir()) / "ctx_hostfield"
        d.mkdir(exist_ok=True)
        so = d / f"f_{key}.so"
        if not so.exists():
            cu = d / f"f_{key}.cpp"
            cu.write_text(src)
            subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-ffp-contract=off", str(cu), "-o", str(so)],
                           check=True)
        self.lib = C.CDLL(str(so))
        self.field = field
        dp = C.POINTER(C.c_double)
        self.lib.eval_rhs.argtypes = [C.c_double, dp, dp, dp, dp, dp]
        if field.D:
            self.lib.eval_tan.argtypes = [C.c_double, dp, dp, dp, dp, dp, dp, dp]
        if getattr(field, "warp_form", False):
            self.lib.eval_warp.argtypes = [C.c_double, dp, dp, dp, dp, dp]

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def rhs(self, t, y, th, c, y0):
        y, th, c, y0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (y, th, c, y0))
        dy = np.zeros(self.field.S)
        self.lib.eval_rhs(float(t), self._p(y), self._p(th), self._p(c), self._p(y0), self._p(dy))
        return dy

    def warp(self, t, y, th, c, y0):
        """The lane-parallel form (generate_warp_form), emulated lane by lane."""
        y, th, c, y0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (y, th, c, y0))
        dy = np.zeros(self.field.S)
        self.lib.eval_warp(float(t), self._p(y), self._p(th), self._p(c), self._p(y0), self._p(dy))
        return dy

    def tan(self, t, y, s, th, c, y0):
        y, s, th, c, y0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (y, s, th, c, y0))
        dy = np.zeros(self.field.S)
        ds = np.zeros(self.field.S * self.field.D)
        self.lib.eval_tan(float(t), self._p(y), self._p(s), self._p(th), self._p(c), self._p(y0),
                          self._p(dy), self._p(ds))
        return dy, ds
