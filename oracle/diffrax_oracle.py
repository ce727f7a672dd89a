"""CPU restatement (numpy) of the reference's batched ODE solve.  PARITY ORACLE.

TEST INFRASTRUCTURE ONLY: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this.
The product (``catalax_b200/``) never does; it has no CPU fallback.

What is restated
----------------
``catalax/tools/simulation.py:236-275`` (``_simulate_system``) calls::

    diffeqsolve(ODETerm(stack), solver(), t0=0, t1=time[-1], dt0=config.dt0, y0=y0,
                args=(parameters, constants, y0), saveat=SaveAt(ts=time),
                stepsize_controller=PIDController(rtol, atol) | ConstantStepSize(),
                max_steps=config.max_steps, throw=config.throw).ys

and ``simulation.py:346-355`` wraps it in ``jax.vmap(..., in_axes)``.  The arithmetic
lives in ``diffrax>=0.7.0,<0.8`` (``pyproject.toml:24``), which is NOT vendored in
``/root/reference`` and cannot be installed here (no jax / diffrax wheels, no network).
This file restates diffrax 0.7's published algorithm:

* explicit Runge-Kutta step with increments ``k_i = f_i * dt`` (FSAL for Tsit5/Dopri5,
  the stored ``f`` is reused after a rejection), stage times ``t0 + c_i*dt``;
* ``y_error = sum_i b_error_i k_i``; NaNs in ``y_error`` become ``inf``;
* ``PIDController`` defaults (pcoeff=0, icoeff=1, dcoeff=0, safety=0.9, factormin=0.2,
  factormax=10, rms norm): ``scaled = rms(y_error / (atol + rtol*max(|y0|,|y1|)))``,
  ``keep = scaled < 1``, ``factor = clip(0.9 * (1/scaled)**(1/order), 1 if keep else 0.2, 10)``,
  ``dt_next = (t1 - t0) * factor``;
* end clipping: if ``tnext > t_end - tol`` (tol 1e-6 f32 / 1e-10 f64) then
  ``tnext = t_end`` on an accepted step, else ``tprev + 0.5*(t_end - tprev)``;
* ``SaveAt(ts)``: after each ACCEPTED step every ``ts[i] <= tnext`` not yet written is
  evaluated with the solver's dense interpolant on ``[tprev, tnext]``;
  unsaved slots stay ``inf`` (``throw=False`` behaviour the MCMC relies on,
  ``catalax/mcmc/mcmc.py:216-219``);
* at most ``max_steps`` loop iterations (accepted + rejected);
* vmap semantics: every lane steps with its own ``dt``/accept decision.

Forward tangents ("discrete sensitivities"): diffrax puts the controller's ``factor``
under ``stop_gradient``, so the derivative of the saved states is the tangent of the
accepted-step sequence *with the step sizes frozen*, dense interpolant included.  That is
the same RK scheme applied to the augmented system with only the first ``n_err``
components in the error norm -- pass ``n_err``.

PARITY STATUS: diffrax cannot be run here.  The restatement is pinned against (a) everything
the reference's tests hold for this path -- the closed forms of
``tests/integration/test_simulation.py:41-195`` and the f32 trajectory fixture
``tests/fixtures/data.npy`` (``tests/test_oracle_golden.py``) -- and (b) 120 trajectories the
reference itself produced with its default config (``examples/datasets/croissant_dataset.zip``,
written by ``examples/GenerateData.ipynb``): their discretisation error is a fingerprint of
diffrax's accepted steps and dense interpolant, which this oracle reproduces up to the one
step-size factor per series that is f32 rounding noise (``factor_override``;
``tests/test_reference_pins.py``).  Bit-level f32 agreement with XLA is not defined.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import tableaus as tb


@dataclass
class OracleResult:
    ys: np.ndarray          # [B, T, N]
    status: np.ndarray      # [B] int32: 0 ok, 1 max_steps reached, 2 non-finite step size
    n_accepted: np.ndarray  # [B] int32
    n_rejected: np.ndarray  # [B] int32
    steps: list | None = None  # optional per-lane accepted (tprev, tnext) history


def _bcast(a, B, dtype):
    a = np.asarray(a, dtype=dtype)
    if a.ndim == 1:
        a = np.broadcast_to(a, (B,) + a.shape)
    return np.ascontiguousarray(a)


def solve(rhs, y0, theta, consts, ts, *, solver="tsit5", rtol=1e-5, atol=1e-5, dt0=0.1,
          max_steps=4096, dtype=np.float64, n_err=None, init_aug=None,
          record_steps=False, forced_steps=None, t0_from_ts=False, factor_override=None):
    """Integrate a batch of trajectories the way the reference does.

    rhs(t[B], Y[B,N], theta[B,P], consts[B,C], y0[B,S]) -> dY[B,N]  (vectorised over B).
    y0 [B,S] or [S]; theta [B,P] or [P]; consts [B,C] or [C]; ts [B,T] or [T].
    n_err : number of leading components that enter the error norm (default all).
    init_aug : optional [B, N-S] initial values of the augmented components (tangents).
    forced_steps : optional list (per lane) of accepted (tprev, tnext) pairs to replay
        instead of running the controller (used to verify tangents by finite differences).
    factor_override : optional [B, K] array; a finite entry [b, n] replaces the controller's
        step-size factor after loop iteration n of lane b (NaN = computed factor).  Used by
        ``tests/test_reference_croissant.py`` to stand in for the step factors that are pure
        f32 rounding noise in the reference's own run.
    """
    dtype = np.dtype(dtype)
    R = dtype.type
    y0 = np.asarray(y0, dtype=dtype)
    B = max(
        y0.shape[0] if y0.ndim == 2 else 1,
        np.asarray(theta).shape[0] if np.asarray(theta).ndim == 2 else 1,
        np.asarray(consts).shape[0] if np.asarray(consts).ndim == 2 else 1,
        np.asarray(ts).shape[0] if np.asarray(ts).ndim == 2 else 1,
    )
    y0 = _bcast(y0, B, dtype)
    theta = _bcast(theta, B, dtype)
    consts = _bcast(consts, B, dtype)
    ts = _bcast(ts, B, dtype)
    S = y0.shape[1]
    T = ts.shape[1]
    if init_aug is not None:
        Y = np.concatenate([y0, _bcast(init_aug, B, dtype)], axis=1)
    else:
        Y = y0.copy()
    N = Y.shape[1]
    n_err = N if n_err is None else n_err

    adaptive = tb.ADAPTIVE[solver]
    order = tb.ORDER[solver]
    if solver == "tsit5":
        A, Bw, E, C = tb.TSIT5_A, tb.TSIT5_B, tb.TSIT5_E, tb.TSIT5_C
    elif solver == "dopri5":
        A, Bw, E, C = tb.DOPRI5_A, tb.DOPRI5_B, tb.DOPRI5_E, tb.DOPRI5_C
    elif solver == "heun":
        A, Bw, E, C = tb.HEUN_A, tb.HEUN_B, None, tb.HEUN_C
    elif solver == "euler":
        A, Bw, E, C = [[]], np.array([1.0]), None, np.array([0.0])
    else:
        raise ValueError(solver)
    fsal = solver in ("tsit5", "dopri5")
    nst = len(C)
    A = [[R(a) for a in row] for row in A]
    Bw = [R(b) for b in Bw]
    E = None if E is None else [R(e) for e in E]
    C = [R(c) for c in C]

    rtol = R(rtol); atol = R(atol); dt0 = R(dt0)
    tol_end = R(1e-10) if dtype == np.float64 else R(1e-6)
    safety, fmin, fmax = R(0.9), R(0.2), R(10.0)
    expo = R(1.0 / order)

    t_end = ts[:, -1].copy()                       # t1 = time[-1]
    # t0 = 0 for mechanistic models (simulation.py:262); the neural solves start at ts[0]
    # (neuralode.py:52, rateflow.py:114); UniversalODE uses 0.0 again (universalode.py:133).
    tprev = ts[:, 0].copy() if t0_from_ts else np.zeros(B, dtype=dtype)
    tnext = np.minimum(tprev + dt0, t_end)
    ys = np.full((B, T, N), np.inf, dtype=dtype)
    save_idx = np.zeros(B, dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    nacc = np.zeros(B, dtype=np.int32)
    nrej = np.zeros(B, dtype=np.int32)
    nsteps = np.zeros(B, dtype=np.int64)
    history = [[] for _ in range(B)] if record_steps else None

    # diffrax: t0 == t1 -> every requested time equals t0; y0 is returned.
    degenerate = ~(tprev < t_end)
    if degenerate.any():
        ys[degenerate] = Y[degenerate][:, None, :]
        save_idx[degenerate] = T

    with np.errstate(all="ignore"):
        f0 = rhs(tprev, Y, theta, consts, y0).astype(dtype) if fsal else None
        forced_ptr = np.zeros(B, dtype=np.int64)
        while True:
            act = np.nonzero((tprev < t_end) & (status == 0))[0]
            if act.size == 0:
                break
            over = nsteps[act] >= max_steps
            if over.any():
                status[act[over]] = 1
                act = act[~over]
                if act.size == 0:
                    continue
            if forced_steps is not None:
                for j in act:
                    tp, tn = forced_steps[j][forced_ptr[j]]
                    tprev[j], tnext[j] = R(tp), R(tn)
                    forced_ptr[j] += 1
            tp = tprev[act]; tn = tnext[act]
            dt = tn - tp
            Ya = Y[act]; th = theta[act]; cc = consts[act]; y0a = y0[act]
            ks = []
            for i in range(nst):
                if i == 0:
                    fi = f0[act] if fsal else rhs(tp, Ya, th, cc, y0a).astype(dtype)
                    yi = Ya
                else:
                    acc = A[i][0] * ks[0]
                    for j in range(1, i):
                        acc = acc + A[i][j] * ks[j]
                    yi = Ya + acc
                    fi = rhs(tp + C[i] * dt, yi, th, cc, y0a).astype(dtype)
                ks.append(fi * dt[:, None])
                f_last, y_last = fi, yi
            if fsal:                                   # SSAL: y1 is the last stage input
                y1 = y_last
            else:
                acc = Bw[0] * ks[0]
                for j in range(1, nst):
                    acc = acc + Bw[j] * ks[j]
                y1 = Ya + acc

            if adaptive and forced_steps is None:
                yerr = E[0] * ks[0]
                for j in range(1, nst):
                    yerr = yerr + E[j] * ks[j]
                yerr = yerr[:, :n_err]
                yerr = np.where(np.isnan(yerr), R(np.inf), yerr)
                y1e = y1[:, :n_err]; y0e = Ya[:, :n_err]
                has_nan = np.isnan(y1e).any(axis=1, keepdims=True)
                y1s = np.where(has_nan, y0e, y1e)
                scale = atol + np.maximum(np.abs(y0e), np.abs(y1s)) * rtol
                ratio = yerr / scale
                scaled = np.sqrt(np.mean(ratio * ratio, axis=1)).astype(dtype)
                keep = scaled < 1
                inv = (R(1.0) / scaled).astype(dtype)
                factor = safety * np.power(inv, expo)
                lo = np.where(keep, R(1.0), fmin)
                factor = np.minimum(np.maximum(factor, lo), fmax).astype(dtype)
                if factor_override is not None:
                    n_it = nsteps[act]
                    inside = n_it < factor_override.shape[1]
                    forced = factor_override[act, np.minimum(n_it, factor_override.shape[1] - 1)]
                    use = inside & np.isfinite(forced)
                    factor = np.where(use, forced, factor).astype(dtype)
                new_dt = (dt * factor).astype(dtype)
            else:
                keep = np.ones(act.size, dtype=bool)
                new_dt = np.full(act.size, dt0, dtype=dtype) if forced_steps is None else dt

            # SaveAt(ts): dense output on accepted steps, ts[i] <= tnext
            for loc in np.nonzero(keep)[0]:
                b = act[loc]
                i = save_idx[b]
                while i < T and ts[b, i] <= tn[loc]:
                    tq = ts[b, i]
                    den = tn[loc] - tp[loc]
                    th_ = R(0.0) if den == 0 else (tq - tp[loc]) / den
                    kk = [k[loc] for k in ks]
                    if solver == "tsit5":
                        w = tb.tsit5_dense_weights(th_)
                        acc = R(w[0]) * kk[0]
                        for j in range(1, 7):
                            acc = acc + R(w[j]) * kk[j]
                        ys[b, i] = Ya[loc] + acc
                    elif solver == "dopri5":
                        acc = R(tb.DOPRI5_CMID[0]) * kk[0]
                        for j in range(1, 7):
                            acc = acc + R(tb.DOPRI5_CMID[j]) * kk[j]
                        ymid = Ya[loc] + acc
                        f0_, f1_ = kk[0], kk[6]
                        y0_, y1_ = Ya[loc], y1[loc]
                        a_ = R(2) * (f1_ - f0_) - R(8) * (y1_ + y0_) + R(16) * ymid
                        b_ = R(5) * f0_ - R(3) * f1_ + R(18) * y0_ + R(14) * y1_ - R(32) * ymid
                        c_ = f1_ - R(4) * f0_ - R(11) * y0_ - R(5) * y1_ + R(16) * ymid
                        ys[b, i] = (((a_ * th_ + b_) * th_ + c_) * th_ + f0_) * th_ + y0_
                    else:                               # LocalLinearInterpolation
                        ys[b, i] = Ya[loc] + th_ * (y1[loc] - Ya[loc])
                    i += 1
                save_idx[b] = i
                if record_steps:
                    history[b].append((float(tp[loc]), float(tn[loc])))

            # accept / reject bookkeeping (diffrax _integrate.py body_fun)
            ka = act[keep]
            Y[ka] = y1[keep]
            if fsal:
                f0[ka] = f_last[keep]
            nacc[ka] += 1
            nrej[act[~keep]] += 1
            nsteps[act] += 1
            nt_prev = np.where(keep, tn, tp)
            nt_prev = np.minimum(nt_prev, t_end[act])
            nt_next = (nt_prev + new_dt).astype(dtype)
            clip = nt_next > t_end[act] - tol_end
            tclip = np.where(keep, t_end[act], nt_prev + R(0.5) * (t_end[act] - nt_prev))
            nt_next = np.where(clip, tclip, nt_next).astype(dtype)
            bad = ~np.isfinite(nt_next) & (nt_prev < t_end[act])
            tprev[act] = nt_prev
            tnext[act] = nt_next
            status[act[bad]] = 2
    return OracleResult(ys, status, nacc, nrej, history)
