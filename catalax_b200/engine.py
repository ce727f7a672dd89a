"""ctypes binding of libctxb200.so (the C ABI in include/ctx_b200.h).

This is the host side of the drop-in seam: the reference builds a jitted JAX closure in
``catalax/tools/simulation.py:160-196`` and calls it at ``catalax/model/model.py:699`` and
``catalax/mcmc/mcmc.py:448``; here the same call lands in CUDA kernels through plain C entry
points.  PyTorch is used only for device memory and the current stream.

There is deliberately NO CPU fallback: if the shared library is missing or no GPU is visible,
every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / "libctxb200.so"
_lib = None

F32, F64 = 0, 1
SOLVER_IDS = {"tsit5": 0, "dopri5": 1, "euler": 2, "heun": 3}
LIKELIHOOD_IDS = {"normal": 0, "softlaplace": 1}


class EngineError(RuntimeError):
    """Raised for every failure of the native engine (carries the library's message)."""


class ctx_model_desc(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("n_params", C.c_int32), ("n_consts", C.c_int32),
                ("n_dirs", C.c_int32), ("n_theta_dirs", C.c_int32), ("dtype", C.c_int32)]


class ctx_solve_cfg(C.Structure):
    _fields_ = [("solver", C.c_int32), ("in_axes", C.c_int32 * 4), ("batch", C.c_int64),
                ("n_times", C.c_int32), ("max_steps", C.c_int32), ("rtol", C.c_double),
                ("atol", C.c_double), ("dt0", C.c_double), ("kernel", C.c_int32),
                ("t0_from_ts", C.c_int32), ("inner", C.c_int32), ("reserved", C.c_int32),
                ("order", C.c_void_p)]


class ctx_adjoint_cfg(C.Structure):
    _fields_ = [("batch", C.c_int64), ("n_times", C.c_int32), ("max_steps", C.c_int32),
                ("cap", C.c_int32), ("ts_per_row", C.c_int32), ("rtol", C.c_double),
                ("atol", C.c_double), ("dt0", C.c_double), ("solver", C.c_int32),
                ("reserved", C.c_int32)]


class ctx_xla_solve_desc(C.Structure):
    """Opaque bytes of the ctx_xla_solve / ctx_xla_solve_sens custom calls (include/ctx_b200.h)."""
    _fields_ = [("abi", C.c_int32), ("reserved", C.c_int32), ("model", C.c_uint64),
                ("ws_bytes", C.c_uint64), ("cfg", ctx_solve_cfg)]


class ctx_loglik_cfg(C.Structure):
    _fields_ = [("solver", C.c_int32), ("likelihood", C.c_int32), ("n_chains", C.c_int64),
                ("n_series", C.c_int32), ("n_times", C.c_int32), ("max_steps", C.c_int32),
                ("kernel", C.c_int32), ("rtol", C.c_double), ("atol", C.c_double),
                ("dt0", C.c_double), ("order", C.c_void_p)]


class ctx_xla_loglik_desc(C.Structure):
    """Opaque bytes of the ctx_xla_loglik_grad custom call."""
    _fields_ = [("abi", C.c_int32), ("has_mask", C.c_int32), ("model", C.c_uint64),
                ("ws_bytes", C.c_uint64), ("cfg", ctx_loglik_cfg)]


ABI_VERSION = 4


def lib() -> C.CDLL:
    """Load libctxb200.so (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise EngineError(
            f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "-- catalax_b200 has no CPU fallback")
    # The library carries the hash of the sources it was built from.  A library that is older
    # than the tree (e.g. after `git pull`) would run new Python / code generation against stale
    # kernels and stale embedded NVRTC sources: rebuild it when a compiler is here, refuse it
    # otherwise.  CTX_B200_SKIP_STAMP_CHECK=1 loads it regardless (developer override).
    from . import _build

    if not os.environ.get("CTX_B200_SKIP_STAMP_CHECK") and _build.lib_stamp() != _build._stamp():
        if os.path.exists(_build.NVCC):
            _build.build()
        else:
            raise EngineError(
                f"{_LIB_PATH} was built from other sources than this tree (stamp "
                f"{_build.lib_stamp()!r}) and nvcc is not available to rebuild it")
    L = C.CDLL(str(_LIB_PATH))
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    L.ctx_abi_version.restype = C.c_int
    L.ctx_device_count.restype = C.c_int
    L.ctx_last_error.restype = C.c_size_t
    L.ctx_last_error.argtypes = [C.c_char_p, C.c_size_t]
    L.ctx_launch_count.restype = C.c_int64
    L.ctx_model_compile.restype = C.c_int
    L.ctx_model_compile.argtypes = [C.c_char_p, C.POINTER(ctx_model_desc), C.POINTER(vp)]
    L.ctx_model_free.restype = None
    L.ctx_model_free.argtypes = [vp]
    L.ctx_model_cubin.restype = C.c_int
    L.ctx_model_cubin.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.ctx_workspace_bytes.restype = C.c_size_t
    L.ctx_workspace_bytes.argtypes = [vp, C.POINTER(ctx_solve_cfg)]
    L.ctx_solve.restype = C.c_int
    L.ctx_solve.argtypes = [vp, C.POINTER(ctx_solve_cfg), vp, vp, vp, vp, vp, vp, vp, vp, vp,
                            C.c_size_t, vp]
    L.ctx_solve_sens.restype = C.c_int
    L.ctx_solve_sens.argtypes = [vp, C.POINTER(ctx_solve_cfg), vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                 vp, C.c_size_t, vp]
    L.ctx_pointwise_loglik.restype = C.c_int
    L.ctx_pointwise_loglik.argtypes = [vp, C.POINTER(ctx_solve_cfg), C.c_int32] + [vp] * 11 + [C.c_size_t, vp]
    L.ctx_loglik_workspace_bytes.restype = C.c_size_t
    L.ctx_loglik_workspace_bytes.argtypes = [vp, C.POINTER(ctx_loglik_cfg)]
    L.ctx_loglik_grad.restype = C.c_int
    L.ctx_loglik_grad.argtypes = [vp, C.POINTER(ctx_loglik_cfg)] + [vp] * 11 + [C.c_size_t, vp]
    L.ctx_measure_fma_peak.restype = C.c_int
    L.ctx_measure_fma_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    L.ctx_rates.restype = C.c_int
    L.ctx_rates.argtypes = [vp, C.c_int64, vp, C.c_int32, vp, vp, C.c_int32, vp, C.c_int32, vp,
                            C.c_int32, vp, vp]
    L.ctx_rate_loglik_grad.restype = C.c_int
    L.ctx_rate_loglik_grad.argtypes = [vp, C.c_int32, C.c_int64, C.c_int32, C.c_int32] + [vp] * 12
    L.ctx_mlp_adjoint_workspace_bytes.restype = C.c_size_t
    L.ctx_mlp_adjoint_workspace_bytes.argtypes = [vp, C.POINTER(ctx_adjoint_cfg)]
    L.ctx_mlp_adjoint_forward.restype = C.c_int
    L.ctx_mlp_adjoint_forward.argtypes = [vp, C.POINTER(ctx_adjoint_cfg)] + [vp] * 8 + [C.c_size_t, vp]
    L.ctx_mlp_adjoint_backward.restype = C.c_int
    L.ctx_mlp_adjoint_backward.argtypes = [vp, C.POINTER(ctx_adjoint_cfg)] + [vp] * 8 + [C.c_size_t, vp]
    L.ctx_solve_host.restype = C.c_int
    L.ctx_solve_host.argtypes = [vp, C.POINTER(ctx_solve_cfg), vp, vp, vp, vp, vp, vp, vp, vp]
    L.ctx_loglik_grad_host.restype = C.c_int
    L.ctx_loglik_grad_host.argtypes = [vp, C.POINTER(ctx_loglik_cfg)] + [vp] * 10
    L.ctx_build_stamp.restype = C.c_char_p
    L.ctx_p2p_create.restype = C.c_int
    L.ctx_p2p_create.argtypes = [C.c_int32, C.c_int32, C.c_size_t, C.POINTER(vp), C.c_char_p]
    L.ctx_p2p_connect.restype = C.c_int
    L.ctx_p2p_connect.argtypes = [vp, C.c_char_p]
    L.ctx_p2p_allreduce.restype = C.c_int
    L.ctx_p2p_allreduce.argtypes = [vp, C.c_int32, C.c_size_t, vp, vp]
    L.ctx_p2p_free.restype = None
    L.ctx_p2p_free.argtypes = [vp]
    L.ctx_p2p_last_error.restype = C.c_size_t
    L.ctx_p2p_last_error.argtypes = [C.c_char_p, C.c_size_t]
    L.ctx_last_solve_kernel.restype = C.c_char_p
    for name in ("ctx_xla_solve", "ctx_xla_solve_sens", "ctx_xla_loglik_grad"):
        fn = getattr(L, name)
        fn.restype = None
        fn.argtypes = [vp, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.ctx_xla_desc_bytes.restype = C.c_size_t
    L.ctx_xla_desc_bytes.argtypes = [C.c_int32]
    L.ctx_xla_last_status.restype = C.c_int
    L.ctx_xla_last_error.restype = C.c_size_t
    L.ctx_xla_last_error.argtypes = [C.c_char_p, C.c_size_t]
    _lib = L
    return L


def last_error() -> str:
    buf = C.create_string_buffer(1 << 16)
    lib().ctx_last_error(buf, len(buf))
    return buf.value.decode(errors="replace")


def _check(rc: int) -> None:
    if rc != 0:
        raise EngineError(f"ctx error {rc}: {last_error()}")


def xla_check() -> None:
    """Raise if an XLA custom-call entry point failed since the last check (that ABI has no
    status channel: failures are parked in a process-wide slot)."""
    rc = int(lib().ctx_xla_last_status())
    if rc != 0:
        buf = C.create_string_buffer(1 << 16)
        lib().ctx_xla_last_error(buf, len(buf))
        raise EngineError(f"ctx error {rc}: {buf.value.decode(errors='replace')}")


def device_count() -> int:
    return int(lib().ctx_device_count())


def measure_fma_peak(dtype: str = "float32") -> float:
    """Measured FFMA (float32) / DFMA (float64) peak of the current GPU in TFLOP/s."""
    v = C.c_double()
    _check(lib().ctx_measure_fma_peak(F64 if np.dtype(dtype) == np.float64 else F32, C.byref(v)))
    return float(v.value)


def launch_count() -> int:
    return int(lib().ctx_launch_count())


def last_solve_kernel() -> str:
    """Kernel the last solve of this thread launched (what ``kernel=0`` resolved to)."""
    return lib().ctx_last_solve_kernel().decode()


def cost_order(n_accepted, n_rejected, heavy_fraction=None):
    """Scheduling hint ``order=`` for the next solve of similar inputs, from the step attempts every
    row needed in a previous one: the ``heavy_fraction`` most expensive rows first (descending
    cost), every other row behind them IN INDEX ORDER.  One launch is bound by the critical path
    of its longest trajectories, so only those need to start early; a full sort would also
    concentrate the cheap rows (refill-bound) and scatter the memory accesses of the bulk
    (measured: DESIGN.md section 4).  int32 permutation of the rows."""
    import torch

    cost = n_accepted.to(torch.int64) + n_rejected.to(torch.int64)
    B = cost.numel()
    if heavy_fraction is None:
        heavy_fraction = float(os.environ.get("CTX_B200_COST_ORDER_FRAC", "0.0625"))
    k = int(min(B, max(1, round(B * heavy_fraction))))
    if k >= B:
        return torch.argsort(cost, descending=True, stable=True).to(torch.int32).contiguous()
    top = torch.topk(cost, k, sorted=True).indices
    keep = torch.ones(B, dtype=torch.bool, device=cost.device)
    keep[top] = False
    rest = torch.nonzero(keep).reshape(-1)
    return torch.cat([top, rest]).to(torch.int32).contiguous()


@dataclass
class SolveOutput:
    ys: "object"            # torch.Tensor [B,T,S]
    status: "object"        # torch.int32 [B]
    n_accepted: "object"
    n_rejected: "object"
    sens: Optional["object"] = None  # [B,T,S,D]

    def cost_order(self):
        return cost_order(self.n_accepted, self.n_rejected)


class CompiledModel:
    """One code-generated vector field bound to the native engine."""

    def __init__(self, field, dtype: str = "float32"):
        """field: catalax_b200.tools.codegen.GeneratedField."""
        self.field = field
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype("float32"), np.dtype("float64")):
            raise ValueError("dtype must be float32 or float64")
        self.desc = ctx_model_desc(field.S, field.P, field.C, field.D, field.n_theta_dirs,
                                   F64 if self.dtype == np.float64 else F32)
        self._h = C.c_void_p()
        _check(lib().ctx_model_compile(field.source.encode(), C.byref(self.desc), C.byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                lib().ctx_model_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ build-time helpers
    def cubin(self, solver: str = "tsit5", tangents: bool = False) -> bytes:
        """NVRTC-compile one kernel variant (works without a GPU) and return the cubin."""
        ptr, n = C.c_void_p(), C.c_size_t()
        _check(lib().ctx_model_cubin(self._h, SOLVER_IDS[solver], int(tangents), C.byref(ptr),
                                     C.byref(n)))
        return C.string_at(ptr, n.value)

    # ------------------------------------------------------------------ compute
    def make_cfg(self, *, solver, in_axes, batch, n_times, max_steps, rtol, atol, dt0, kernel=0,
                 t0_from_ts=False, inner=0):
        cfg = ctx_solve_cfg()
        cfg.solver = SOLVER_IDS[solver]
        for i, a in enumerate(in_axes):
            cfg.in_axes[i] = 0 if a == 0 else -1
        cfg.batch, cfg.n_times, cfg.max_steps = int(batch), int(n_times), int(max_steps)
        cfg.rtol, cfg.atol, cfg.dt0 = float(rtol), float(atol), float(dt0)
        cfg.kernel = int(kernel)
        cfg.t0_from_ts = int(bool(t0_from_ts))
        cfg.inner = int(inner)
        return cfg

    def solve(self, y0, theta, consts, ts, *, solver="tsit5", in_axes=(0, None, 0, None),
              rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096, tangents=False, kernel=0,
              out=None, t0_from_ts=False, inner=0, order=None, probe=False):
        """Device-resident solve.  All inputs are torch CUDA tensors of this model's dtype.

        probe=True (opt-in, no ``order`` given, 2^16..2^22 rows): the library first classifies the
        rows with its scheduling probe (csrc/ctx_sched.cu: a solve at 100 x rtol capped at 32 step
        attempts) and starts the rows that hit the cap first.  Pays only where a looser tolerance
        keeps the long rows long; on the benchmark workload it does not.

        order: optional scheduling hint -- an int32 CUDA tensor holding a permutation of the rows,
        e.g. ``previous_output.cost_order()`` ("most expensive first" from the step counts of a
        previous solve of similar inputs).  The persistent kernels hand out rows in that order;
        results do not depend on it.

        inner=M > 0: two-level batch -- theta is [D,P] (one row per posterior draw), y0 / consts /
        ts have M rows (one per series) and ys is [D*M, T, S] (predictive.py:150-154)."""
        import torch

        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        dev = y0.device
        if dev.type != "cuda":
            raise EngineError("solve() needs CUDA tensors: catalax_b200 has no CPU fallback")
        S, P, Cn, D = self.field.S, self.field.P, self.field.C, self.field.D

        def prep(x, width, ax):
            x = x.to(device=dev, dtype=tdt).contiguous()
            want = 2 if ax == 0 else 1
            if x.dim() != want or (x.shape[-1] != width and width >= 0):
                raise ValueError(f"expected a {want}-D array with last dim {width}, got {tuple(x.shape)}")
            return x

        if inner:
            if in_axes[1] != 0 or in_axes[0] != 0:
                raise ValueError("two-level batch needs batched y0 and parameters")
            B = theta.shape[0] * int(inner)
            # y0 / constants / times are indexed by the series m < inner: anything shorter would
            # be read out of bounds on the device
            for name, arr, ax in (("y0", y0, in_axes[0]), ("consts", consts, in_axes[2]),
                                  ("ts", ts, in_axes[3])):
                if ax == 0 and arr.shape[0] != int(inner):
                    raise ValueError(f"two-level batch: {name} must have inner={int(inner)} rows, "
                                     f"got {arr.shape[0]}")
        else:
            Bs = [a.shape[0] for a, ax in ((y0, in_axes[0]), (theta, in_axes[1]),
                                           (consts, in_axes[2]), (ts, in_axes[3])) if ax == 0]
            B = Bs[0] if Bs else 1
            if any(b != B for b in Bs):
                raise ValueError(f"inconsistent batch sizes {Bs}")
        T = ts.shape[-1]
        # kernel 3 (tensor-core MLP) takes its own parameter pack [aux | UMMA weights]
        y0 = prep(y0, S, in_axes[0]); theta = prep(theta, -1 if kernel == 3 else P, in_axes[1])
        consts = prep(consts, Cn, in_axes[2]); ts = prep(ts, T, in_axes[3])
        cfg = self.make_cfg(solver=solver, in_axes=in_axes, batch=B, n_times=T,
                            max_steps=max_steps, rtol=rtol, atol=atol, dt0=dt0, kernel=kernel,
                            t0_from_ts=t0_from_ts, inner=inner)
        if out is not None:
            if (tuple(out.shape) != (B, T, S) or out.dtype != tdt or out.device != dev
                    or not out.is_contiguous()):
                raise ValueError(f"out must be a contiguous {tdt} tensor of shape {(B, T, S)} on {dev}, "
                                 f"got {tuple(out.shape)} {out.dtype} on {out.device}")
        if order is not None:
            if (order.dtype != torch.int32 or order.device != dev or tuple(order.shape) != (B,)
                    or not order.is_contiguous()):
                raise ValueError(f"order must be a contiguous int32 tensor of shape ({B},) on {dev}")
            cfg.order = order.data_ptr()
        if probe:
            cfg.reserved |= 2
        with torch.cuda.device(dev):
            ys = out if out is not None else torch.empty((B, T, S), dtype=tdt, device=dev)
            status = torch.empty(B, dtype=torch.int32, device=dev)
            nacc = torch.empty(B, dtype=torch.int32, device=dev)
            nrej = torch.empty(B, dtype=torch.int32, device=dev)
            wsb = int(lib().ctx_workspace_bytes(self._h, C.byref(cfg)))
            ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            p = lambda t: C.c_void_p(t.data_ptr()) if t.numel() else C.c_void_p(0)
            if tangents:
                sens = torch.empty((B, T, S, D), dtype=tdt, device=dev)
                _check(lib().ctx_solve_sens(self._h, C.byref(cfg), p(y0), p(theta), p(consts), p(ts),
                                            p(ys), p(sens), p(status), p(nacc), p(nrej), p(ws), wsb,
                                            C.c_void_p(stream)))
            else:
                sens = None
                _check(lib().ctx_solve(self._h, C.byref(cfg), p(y0), p(theta), p(consts), p(ts), p(ys),
                                       p(status), p(nacc), p(nrej), p(ws), wsb, C.c_void_p(stream)))
        return SolveOutput(ys, status, nacc, nrej, sens)

    def pointwise_loglik(self, y0, theta, consts, ts, data, sigma, *, likelihood="normal",
                         solver="tsit5", in_axes=(0, 0, 0, 0), rtol=1e-5, atol=1e-5, dt0=0.1,
                         max_steps=4096, inner=0, out=None):
        """Fused solve + point-wise log-likelihood (ctx_pointwise_loglik): like ``solve`` but the
        dense-output pass writes ``log p(data | loc = y(t), |sigma|)`` [B,T,S] instead of the
        states.  Two-level batch (``inner=M``): theta [D,P] and sigma [D,S] per draw, y0 / consts /
        ts / data [M,...] per series -> [D*M, T, S] (loo.py:180-193)."""
        import torch

        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        dev = y0.device
        if dev.type != "cuda":
            raise EngineError("pointwise_loglik() needs CUDA tensors: catalax_b200 has no CPU fallback")
        S, P, Cn = self.field.S, self.field.P, self.field.C
        f = lambda x: x.to(device=dev, dtype=tdt).contiguous()
        y0, theta, consts, ts, data, sigma = map(f, (y0, theta, consts, ts, data, sigma))
        T = ts.shape[-1]
        if inner:
            M = int(inner)
            D = theta.shape[0]
            B = D * M
            if (y0.shape != (M, S) or consts.shape != (M, Cn) or ts.shape != (M, T)
                    or data.shape != (M, T, S) or sigma.shape != (D, S) or theta.shape != (D, P)):
                raise ValueError("pointwise_loglik: two-level batch wants y0 [M,S], consts [M,C], "
                                 "ts [M,T], data [M,T,S], theta [D,P], sigma [D,S]")
            in_axes = (0, 0, 0, 0)
        else:
            B = y0.shape[0]
            if data.shape != (B, T, S) or sigma.shape != (B, S):
                raise ValueError("pointwise_loglik: data must be [B,T,S] and sigma [B,S]")
        if not bool(torch.isfinite(data).all()):
            raise ValueError("data must be finite (use 0 at masked points and mask the result)")
        cfg = self.make_cfg(solver=solver, in_axes=in_axes, batch=B, n_times=T, max_steps=max_steps,
                            rtol=rtol, atol=atol, dt0=dt0, kernel=0, inner=inner)
        if out is not None and (tuple(out.shape) != (B, T, S) or out.dtype != tdt or out.device != dev
                                or not out.is_contiguous()):
            raise ValueError(f"out must be a contiguous {tdt} tensor of shape {(B, T, S)} on {dev}")
        with torch.cuda.device(dev):
            ll = out if out is not None else torch.empty((B, T, S), dtype=tdt, device=dev)
            cnt = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(3)]
            wsb = int(lib().ctx_workspace_bytes(self._h, C.byref(cfg)))
            ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            p = lambda t: C.c_void_p(t.data_ptr()) if t.numel() else C.c_void_p(0)
            _check(lib().ctx_pointwise_loglik(self._h, C.byref(cfg), LIKELIHOOD_IDS[likelihood], p(y0),
                                              p(theta), p(consts), p(ts), p(data), p(sigma), p(ll),
                                              p(cnt[0]), p(cnt[1]), p(cnt[2]), p(ws), wsb,
                                              C.c_void_p(stream)))
        return SolveOutput(ll, cnt[0], cnt[1], cnt[2])

    def solve_host(self, y0, theta, consts, ts, *, solver="tsit5", in_axes=(0, None, 0, None),
                   rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096, kernel=0, out=None):
        """Host-buffer solve through ctx_solve_host: numpy in, numpy out, copies inside."""
        dt = self.dtype
        S, P, Cn = self.field.S, self.field.P, self.field.C
        y0 = np.ascontiguousarray(y0, dtype=dt); theta = np.ascontiguousarray(theta, dtype=dt)
        consts = np.ascontiguousarray(consts, dtype=dt); ts = np.ascontiguousarray(ts, dtype=dt)
        Bs = [a.shape[0] for a, ax in ((y0, in_axes[0]), (theta, in_axes[1]),
                                       (consts, in_axes[2]), (ts, in_axes[3])) if ax == 0]
        B = Bs[0] if Bs else 1
        T = ts.shape[-1]
        cfg = self.make_cfg(solver=solver, in_axes=in_axes, batch=B, n_times=T,
                            max_steps=max_steps, rtol=rtol, atol=atol, dt0=dt0, kernel=kernel)
        if any(b != B for b in Bs):
            raise ValueError(f"inconsistent batch sizes {Bs}")
        for name, arr, ax, width in (("y0", y0, in_axes[0], S), ("theta", theta, in_axes[1], P),
                                     ("consts", consts, in_axes[2], Cn), ("ts", ts, in_axes[3], T)):
            if arr.ndim != (2 if ax == 0 else 1) or arr.shape[-1] != width:
                raise ValueError(f"{name}: expected last dim {width} and "
                                 f"{2 if ax == 0 else 1} dims, got {arr.shape}")
        if out is not None and (out.shape != (B, T, S) or out.dtype != dt
                                or not out.flags["C_CONTIGUOUS"]):
            raise ValueError(f"out must be a C-contiguous {dt} array of shape {(B, T, S)}")
        ys = out if out is not None else np.empty((B, T, S), dtype=dt)
        status = np.empty(B, dtype=np.int32); nacc = np.empty(B, dtype=np.int32)
        nrej = np.empty(B, dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        _check(lib().ctx_solve_host(self._h, C.byref(cfg), vp(y0), vp(theta), vp(consts), vp(ts),
                                    vp(ys), vp(status), vp(nacc), vp(nrej)))
        return SolveOutput(ys, status, nacc, nrej)

    def solve_xla(self, y0, theta, consts, ts, *, solver="tsit5", in_axes=(0, None, 0, None),
                  rtol=1e-5, atol=1e-5, dt0=0.1, max_steps=4096, tangents=False, kernel=0,
                  t0_from_ts=False, inner=0):
        """The same solve through the XLA custom-call entry point (ctx_xla_solve[_sens]): a
        `void* buffers[]` of operands then results and the packed descriptor as opaque bytes --
        exactly what XLA's GPU runtime passes (include/ctx_b200.h)."""
        import torch

        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        dev = y0.device
        if dev.type != "cuda":
            raise EngineError("solve_xla() needs CUDA tensors: catalax_b200 has no CPU fallback")
        S, D = self.field.S, self.field.D
        f = lambda x: x.to(device=dev, dtype=tdt).contiguous()
        y0, theta, consts, ts = f(y0), f(theta), f(consts), f(ts)
        if inner:
            B = theta.shape[0] * int(inner)
        else:
            Bs = [a.shape[0] for a, ax in ((y0, in_axes[0]), (theta, in_axes[1]),
                                           (consts, in_axes[2]), (ts, in_axes[3])) if ax == 0]
            B = Bs[0] if Bs else 1
        T = ts.shape[-1]
        d = ctx_xla_solve_desc()
        d.abi, d.model = ABI_VERSION, self._h.value
        d.cfg = self.make_cfg(solver=solver, in_axes=in_axes, batch=B, n_times=T,
                              max_steps=max_steps, rtol=rtol, atol=atol, dt0=dt0, kernel=kernel,
                              t0_from_ts=t0_from_ts, inner=inner)
        d.ws_bytes = int(lib().ctx_workspace_bytes(self._h, C.byref(d.cfg)))
        with torch.cuda.device(dev):
            ys = torch.empty((B, T, S), dtype=tdt, device=dev)
            sens = torch.empty((B, T, S, D), dtype=tdt, device=dev) if tangents else None
            cnt = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(3)]
            ws = torch.empty(max(int(d.ws_bytes), 1), dtype=torch.uint8, device=dev)
            bufs = [y0, theta, consts, ts, ys] + ([sens] if tangents else []) + cnt + [ws]
            arr = (C.c_void_p * len(bufs))(*[b.data_ptr() if b.numel() else 0 for b in bufs])
            stream = torch.cuda.current_stream(dev).cuda_stream
            opaque = bytes(d)
            fn = lib().ctx_xla_solve_sens if tangents else lib().ctx_xla_solve
            fn(C.c_void_p(stream), arr, opaque, len(opaque))
        xla_check()
        return SolveOutput(ys, cnt[0], cnt[1], cnt[2], sens)

    def loglik_grad_xla(self, y0s, theta, consts, ts, data, sigma, *, mask=None,
                        likelihood="softlaplace", solver="tsit5", rtol=1e-5, atol=1e-5, dt0=0.1,
                        max_steps=4096, kernel=0):
        """loglik_grad through the XLA custom-call entry point (ctx_xla_loglik_grad)."""
        import torch

        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        dev = theta.device
        S, P, D = self.field.S, self.field.P, self.field.D
        f = lambda x: x.to(device=dev, dtype=tdt).contiguous()
        y0s, theta, consts, ts, data, sigma = map(f, (y0s, theta, consts, ts, data, sigma))
        M, T = ts.shape
        CH = theta.shape[0]
        d = ctx_xla_loglik_desc()
        d.abi, d.model, d.has_mask = ABI_VERSION, self._h.value, int(mask is not None)
        d.cfg.solver, d.cfg.likelihood = SOLVER_IDS[solver], LIKELIHOOD_IDS[likelihood]
        d.cfg.n_chains, d.cfg.n_series, d.cfg.n_times, d.cfg.max_steps = CH, M, T, int(max_steps)
        d.cfg.rtol, d.cfg.atol, d.cfg.dt0, d.cfg.kernel = float(rtol), float(atol), float(dt0), int(kernel)
        d.ws_bytes = int(lib().ctx_loglik_workspace_bytes(self._h, C.byref(d.cfg)))
        with torch.cuda.device(dev):
            mk = (mask.to(device=dev, dtype=torch.uint8).contiguous() if mask is not None
                  else torch.zeros(1, dtype=torch.uint8, device=dev))   # XLA has no null operands
            out = torch.empty((CH, 1 + P + S), dtype=tdt, device=dev)
            partial = torch.empty((CH * M, 1 + P + S + (D - P)), dtype=tdt, device=dev)
            status = torch.empty(CH * M, dtype=torch.int32, device=dev)
            ws = torch.empty(max(int(d.ws_bytes), 1), dtype=torch.uint8, device=dev)
            bufs = [y0s, theta, consts, ts, data, mk, sigma, out, partial, status, ws]
            arr = (C.c_void_p * len(bufs))(*[b.data_ptr() if b.numel() else 0 for b in bufs])
            stream = torch.cuda.current_stream(dev).cuda_stream
            opaque = bytes(d)
            lib().ctx_xla_loglik_grad(C.c_void_p(stream), arr, opaque, len(opaque))
        xla_check()
        return out, partial, status

    def loglik_grad_host(self, y0s, theta, consts, ts, data, sigma, *, mask=None,
                         likelihood="softlaplace", solver="tsit5", rtol=1e-5, atol=1e-5, dt0=0.1,
                         max_steps=4096, kernel=0, out=None, want_partial=False):
        """Host-buffer log-likelihood + gradient (ctx_loglik_grad_host): numpy in, numpy out; the
        H2D copies of every operand and the D2H copy of `out` happen inside the call."""
        dt = self.dtype
        S, P, Cn, D = self.field.S, self.field.P, self.field.C, self.field.D
        if self.field.n_theta_dirs != P:
            raise EngineError("model must be generated with sens_theta=True")
        g = lambda x: np.ascontiguousarray(x, dtype=dt)
        y0s, theta, consts, ts, data, sigma = map(g, (y0s, theta, consts, ts, data, sigma))
        M, T = ts.shape
        CH = theta.shape[0]
        if (y0s.shape != (M, S) or theta.shape != (CH, P) or data.shape != (M, T, S)
                or sigma.shape != (CH, S) or consts.shape != (M, Cn)):
            raise ValueError("loglik_grad_host: inconsistent operand shapes")
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            if mask.shape != (M, T, S):
                raise ValueError("mask must be [M,T,S]")
        cfg = ctx_loglik_cfg()
        cfg.solver, cfg.likelihood = SOLVER_IDS[solver], LIKELIHOOD_IDS[likelihood]
        cfg.n_chains, cfg.n_series, cfg.n_times, cfg.max_steps = CH, M, T, int(max_steps)
        cfg.rtol, cfg.atol, cfg.dt0, cfg.kernel = float(rtol), float(atol), float(dt0), int(kernel)
        if out is None:
            out = np.empty((CH, 1 + P + S), dtype=dt)
        elif out.shape != (CH, 1 + P + S) or out.dtype != dt or not out.flags["C_CONTIGUOUS"]:
            raise ValueError(f"out must be a C-contiguous {dt} array of shape {(CH, 1 + P + S)}")
        partial = np.empty((CH * M, 1 + P + S + (D - P)), dtype=dt) if want_partial else None
        status = np.empty(CH * M, dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if (a is not None and a.size) else C.c_void_p(0)
        _check(lib().ctx_loglik_grad_host(self._h, C.byref(cfg), vp(y0s), vp(theta), vp(consts),
                                          vp(ts), vp(data), vp(mask), vp(sigma), vp(out),
                                          vp(partial), vp(status)))
        return out, partial, status

    def loglik_grad(self, y0s, theta, consts, ts, data, sigma, *, mask=None,
                    likelihood="softlaplace", solver="tsit5", rtol=1e-5, atol=1e-5, dt0=0.1,
                    max_steps=4096, kernel=0, order=None):
        """Fused log-likelihood + gradient for CH chains over M series (ctx_loglik_grad).
        order: optional int32 permutation of the CH*M trajectories (scheduling hint, e.g.
        ``engine.cost_order(*model.last_loglik_counts)`` of the previous evaluation).
        kernel: 0 automatic | 1 per-lane likelihood terms | 2 warp-cooperative likelihood terms.

        y0s [M,S], theta [CH,P], consts [M,C], ts [M,T], data [M,T,S], sigma [CH,S],
        mask [M,T,S] bool or None (= isfinite(data)).  Returns (out [CH, 1+P+S],
        partial [CH*M, 1+P+S+D_y0], status [CH*M]) as torch CUDA tensors.
        """
        import torch

        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        dev = theta.device
        if dev.type != "cuda":
            raise EngineError("loglik_grad() needs CUDA tensors: catalax_b200 has no CPU fallback")
        S, P, Cn, D = self.field.S, self.field.P, self.field.C, self.field.D
        if self.field.n_theta_dirs != P:
            raise EngineError("model must be generated with sens_theta=True")
        f = lambda x: x.to(device=dev, dtype=tdt).contiguous()
        y0s, theta, consts, ts, data, sigma = map(f, (y0s, theta, consts, ts, data, sigma))
        M, T = ts.shape
        CH = theta.shape[0]
        assert y0s.shape == (M, S) and theta.shape == (CH, P) and data.shape == (M, T, S)
        assert sigma.shape == (CH, S) and consts.shape == (M, Cn)
        if mask is not None:
            mask = mask.to(device=dev, dtype=torch.uint8).contiguous()
            assert mask.shape == (M, T, S)
        cfg = ctx_loglik_cfg()
        cfg.solver, cfg.likelihood = SOLVER_IDS[solver], LIKELIHOOD_IDS[likelihood]
        cfg.n_chains, cfg.n_series, cfg.n_times, cfg.max_steps = CH, M, T, int(max_steps)
        cfg.rtol, cfg.atol, cfg.dt0 = float(rtol), float(atol), float(dt0)
        cfg.kernel = int(kernel)
        if order is not None:
            if (order.dtype != torch.int32 or order.device != dev or tuple(order.shape) != (CH * M,)
                    or not order.is_contiguous()):
                raise ValueError(f"order must be a contiguous int32 tensor of shape ({CH * M},) on {dev}")
            cfg.order = order.data_ptr()
        kp = 1 + P + S + (D - P)
        with torch.cuda.device(dev):
            out = torch.empty((CH, 1 + P + S), dtype=tdt, device=dev)
            partial = torch.empty((CH * M, kp), dtype=tdt, device=dev)
            status = torch.empty(CH * M, dtype=torch.int32, device=dev)
            wsb = int(lib().ctx_loglik_workspace_bytes(self._h, C.byref(cfg)))
            ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            p = lambda t: C.c_void_p(t.data_ptr()) if (t is not None and t.numel()) else C.c_void_p(0)
            _check(lib().ctx_loglik_grad(self._h, C.byref(cfg), p(y0s), p(theta), p(consts), p(ts),
                                         p(data), p(mask), p(sigma), p(out), p(partial), p(status),
                                         p(ws), wsb, C.c_void_p(stream)))
            # step counters live in the workspace: [256 B][status slot][n_accepted][n_rejected]
            n = CH * M
            counts = ws[256 + 4 * n: 256 + 12 * n].view(torch.int32)
            self.last_loglik_counts = (counts[:n], counts[n:])
        return out, partial, status

    def rate_loglik_grad(self, t, y, theta, consts, inits, data, jac2, sigma, *, n_time,
                         mask=None, extra2=None, likelihood="softlaplace"):
        """Surrogate (rate-matching) log-likelihood + gradient for CH chains (ctx_rate_loglik_grad).

        t [N], y [N,S] (measured states, grouped per measurement: p = m*n_time + j), theta [CH,P],
        consts [M,C], inits [M,S], data [N,S] (surrogate rates), jac2 [N,S,S] (squared surrogate
        rate-Jacobian), sigma [CH,S], mask [N,S] or None (= isfinite(data)), extra2 [N,S] or None.
        Returns out [CH, 1+P+S] = (log L, d/dtheta, d/dsigma_y) as a torch CUDA tensor."""
        import torch

        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        dev = theta.device
        if dev.type != "cuda":
            raise EngineError("rate_loglik_grad() needs CUDA tensors: catalax_b200 has no CPU fallback")
        S, P, Cn = self.field.S, self.field.P, self.field.C
        if self.field.n_theta_dirs != P or P == 0:
            raise EngineError("model must be generated with sens_theta=True")
        f = lambda x: x.to(device=dev, dtype=tdt).contiguous()
        t, y, theta, consts, inits, data, jac2, sigma = map(
            f, (t, y, theta, consts, inits, data, jac2, sigma))
        N, CH = y.shape[0], theta.shape[0]
        n_time = int(n_time)
        if n_time <= 0 or N % n_time:
            raise ValueError("the N points must be grouped per measurement: N = M * n_time")
        M = N // n_time
        assert t.shape == (N,) and y.shape == (N, S) and data.shape == (N, S)
        assert jac2.shape == (N, S, S) and sigma.shape == (CH, S) and theta.shape == (CH, P)
        assert inits.shape == (M, S) and consts.shape == (M, Cn)
        if mask is not None:
            mask = mask.to(device=dev, dtype=torch.uint8).contiguous()
            assert mask.shape == (N, S)
        if extra2 is not None:
            extra2 = f(extra2)
            assert extra2.shape == (N, S)
        out = torch.empty((CH, 1 + P + S), dtype=tdt, device=dev)
        p = lambda x: C.c_void_p(x.data_ptr()) if (x is not None and x.numel()) else C.c_void_p(0)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _check(lib().ctx_rate_loglik_grad(
                self._h, LIKELIHOOD_IDS[likelihood], CH, N, n_time, p(t), p(y), p(theta), p(consts),
                p(inits), p(data), p(mask), p(jac2), p(extra2), p(sigma), p(out),
                C.c_void_p(stream)))
        return out

    # ------------------------------------------------------------------ NeuralODE training
    def adjoint_forward(self, y0, theta, ts, *, rtol=1e-3, atol=1e-6, dt0=0.1, max_steps=4096,
                        cap=512, solver="tsit5"):
        """Forward solve that records every accepted step (ctx_mlp_adjoint_forward).
        y0 [B,S], theta [P] (flat layout), ts [B,T] or [T].  Returns a context dict holding ys
        [B,T,S], status / step counts and the workspace `adjoint_backward` needs."""
        import torch

        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        dev = y0.device
        if dev.type != "cuda":
            raise EngineError("adjoint_forward() needs CUDA tensors: catalax_b200 has no CPU fallback")
        f = lambda x: x.to(device=dev, dtype=tdt).contiguous()
        y0, theta, ts = f(y0), f(theta), f(ts)
        B, S = y0.shape
        T = ts.shape[-1]
        cfg = ctx_adjoint_cfg()
        cfg.batch, cfg.n_times, cfg.max_steps, cfg.cap = B, T, int(max_steps), int(cap)
        cfg.ts_per_row = 1 if ts.dim() == 2 else 0
        cfg.rtol, cfg.atol, cfg.dt0 = float(rtol), float(atol), float(dt0)
        cfg.solver = SOLVER_IDS[solver]
        p = lambda t: C.c_void_p(t.data_ptr()) if (t is not None and t.numel()) else C.c_void_p(0)
        with torch.cuda.device(dev):
            wsb = int(lib().ctx_mlp_adjoint_workspace_bytes(self._h, C.byref(cfg)))
            if wsb == 0:
                raise EngineError(last_error() or "adjoint workspace query failed")
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            ys = torch.empty((B, T, S), dtype=tdt, device=dev)
            counts = torch.empty((3, B), dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _check(lib().ctx_mlp_adjoint_forward(self._h, C.byref(cfg), p(y0), p(theta), p(ts), p(ys),
                                                 p(counts[0]), p(counts[1]), p(counts[2]), p(ws),
                                                 wsb, C.c_void_p(stream)))
        return dict(cfg=cfg, ws=ws, wsb=wsb, ys=ys, status=counts[0], n_accepted=counts[1],
                    n_rejected=counts[2], theta=theta, ts=ts, y0=y0)

    def adjoint_backward(self, ctx, gys, want_y0: bool = False):
        """d loss / d theta [P] (and d loss / d y0 [B,S]) from gys = d loss / d ys [B,T,S]."""
        import torch

        ys = ctx["ys"]
        gys = gys.to(device=ys.device, dtype=ys.dtype).contiguous()
        assert gys.shape == ys.shape
        p = lambda t: C.c_void_p(t.data_ptr()) if (t is not None and t.numel()) else C.c_void_p(0)
        with torch.cuda.device(ys.device):
            g = torch.empty(ctx["theta"].shape, dtype=ys.dtype, device=ys.device)
            gy0 = torch.empty((ys.shape[0], ys.shape[2]), dtype=ys.dtype, device=ys.device) \
                if want_y0 else None
            stream = torch.cuda.current_stream(ys.device).cuda_stream
            _check(lib().ctx_mlp_adjoint_backward(self._h, C.byref(ctx["cfg"]), p(ctx["y0"]),
                                                  p(ctx["theta"]), p(ctx["ts"]), p(gys),
                                                  p(ctx["status"]), p(g), p(gy0),
                                                  p(ctx["ws"]), ctx["wsb"], C.c_void_p(stream)))
        return (g, gy0) if want_y0 else g

    def rates(self, t, y, theta, consts, y0=None):
        """Right-hand side at N points (ctx_rates): t [N]|scalar, y [N,S], theta [N,P]|[P],
        consts [N,C]|[C], y0 [N,S]|[S]|None.  torch CUDA tensors in, [N,S] out."""
        import torch

        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        dev = y.device
        if dev.type != "cuda":
            raise EngineError("rates() needs CUDA tensors: catalax_b200 has no CPU fallback")
        f = lambda x: x.to(device=dev, dtype=tdt).contiguous()
        y = f(y)
        N = y.shape[0]
        t = f(torch.as_tensor(t)).reshape(-1)
        theta, consts = f(theta), f(consts)
        y0 = None if y0 is None else f(y0)
        st = lambda x, w: 0 if x is None or x.dim() == 1 else w
        out = torch.empty((N, self.field.S), dtype=tdt, device=dev)
        p = lambda x: C.c_void_p(x.data_ptr()) if (x is not None and x.numel()) else C.c_void_p(0)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _check(lib().ctx_rates(self._h, N, p(t), 1 if t.numel() > 1 else 0, p(y), p(theta),
                                   st(theta, theta.shape[-1]), p(consts), st(consts, self.field.C),
                                   p(y0), st(y0, self.field.S), p(out), C.c_void_p(stream)))
        return out
