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

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def rhs(self, t, y, th, c, y0):
        y, th, c, y0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (y, th, c, y0))
        dy = np.zeros(self.field.S)
        self.lib.eval_rhs(float(t), self._p(y), self._p(th), self._p(c), self._p(y0), self._p(dy))
        return dy

    def tan(self, t, y, s, th, c, y0):
        y, s, th, c, y0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (y, s, th, c, y0))
        dy = np.zeros(self.field.S)
        ds = np.zeros(self.field.S * self.field.D)
        self.lib.eval_tan(float(t), self._p(y), self._p(s), self._p(th), self._p(c), self._p(y0),
                          self._p(dy), self._p(ds))
        return dy, ds
