"""CPU restatement of the NeuralODE training gradient.  PARITY ORACLE (test infra).

Follows ``catalax/neural/trainer.py:268-306`` (``eqx.filter_value_and_grad`` of the masked
per-point loss of ``vmap(model)(ti, y0i)``) and ``catalax/neural/neuralode.py:38-59``
(``diffeqsolve(ODETerm(MLP), Tsit5, t0=ts[0], t1=ts[-1], dt0, y0, PIDController, SaveAt(ts))``).
diffrax's default ``RecursiveCheckpointAdjoint`` is reverse-mode differentiation of the discrete
scheme with the controller's step factor under ``stop_gradient``.  Restated here with torch
autograd (float64, CPU): the adaptive Tsit5 loop of ``oracle/diffrax_oracle.py`` written with
torch ops, accept / reject decisions and step sizes computed from detached values, dense output
through the solver's own interpolant.  PARITY UNPINNED against diffrax / equinox themselves
(un-vendored dependencies, ``pyproject.toml:24-26``); cross-checked against finite differences of
the replayed step sequence in ``tests/``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import tableaus as tb

_ACT = {
    "tanh": torch.tanh,
    "softplus": torch.nn.functional.softplus,
    "celu": torch.nn.functional.celu,
    "relu": torch.relu,
    "identity": lambda x: x,
}


def sympy_field_torch(states, parameters, exprs):
    """The mechanistic stack (catalax/tools/stack.py:159-173) as a torch function
    ``mech(t, y [S], theta [P], y0 [S]) -> [S]`` (sympy.lambdify with torch functions)."""
    import sympy as sp

    args = [sp.Symbol("t")] + [sp.Symbol(s) for s in states] + [sp.Symbol(p) for p in parameters] \
        + [sp.Symbol(f"{s}_init") for s in states]
    mods = [{"exp": torch.exp, "log": torch.log, "sqrt": torch.sqrt, "tanh": torch.tanh,
             "sin": torch.sin, "cos": torch.cos, "Abs": torch.abs}, "math"]
    funs = [sp.lambdify(args, sp.sympify(e), modules=mods) for e in exprs]

    def mech(t, y, theta, y0):
        vals = [t] + [y[i] for i in range(len(states))] + [theta[i] for i in range(len(parameters))] \
            + [y0[i] for i in range(len(states))]
        out = [fn(*vals) for fn in funs]
        return torch.stack([o if torch.is_tensor(o) else torch.tensor(float(o), dtype=torch.float64)
                            for o in out])
    return mech


def mlp_field(weights, biases, activation, final_activation, max_time, stoich=None, uode=None):
    """NeuralODE field (neuralode.py:49); with ``stoich`` [S,R]: RateFlowODE's
    ``stoich @ relu(MLP)`` (rateflow.py:87-88, final_activation relu :56); with
    ``uode = (mech, theta_m, alpha_residual, gate_matrix, y0)``: UniversalODE's
    ``mech + softplus(alpha) * sigmoid(10 G y) * MLP`` (universalode.py:84-101)."""
    def f(t, y):
        h = torch.cat([y, (t / max_time).reshape(1)])
        for l, (W, b) in enumerate(zip(weights, biases)):
            h = W @ h
            if b is not None:
                h = h + b
            h = _ACT[final_activation if l == len(weights) - 1 else activation](h)
        if stoich is not None:
            h = stoich @ torch.relu(h)
        if uode is not None:
            mech, theta_m, alpha, gate, y_init = uode
            h = mech(t, y, theta_m, y_init) + torch.nn.functional.softplus(alpha) \
                * torch.sigmoid(10.0 * (gate @ y)) * h
        return h
    return f


def solve_one(f, ts, y0, rtol, atol, dt0, max_steps=4096, t0=None, solver="tsit5"):
    """One trajectory: ts [T] (float64 tensor), y0 [S] -> ys [T,S] (differentiable), n_acc, n_rej.
    t0 = ts[0] (neuralode.py:52) unless given (universalode.py:133: 0.0)."""
    T = ts.shape[0]
    t_end = float(ts[-1])
    tprev = float(ts[0]) if t0 is None else float(t0)
    tnext = min(tprev + dt0, t_end)
    y = y0
    out = [None] * T
    idx = 0
    if not tprev < t_end:
        return torch.stack([y0 for _ in range(T)]), 0, 0
    k1 = f(torch.tensor(tprev, dtype=torch.float64), y)
    nacc = nrej = 0
    A, C, E = (tb.TSIT5_A, tb.TSIT5_C, tb.TSIT5_E) if solver == "tsit5" else \
        (tb.DOPRI5_A, tb.DOPRI5_C, tb.DOPRI5_E)
    dense_weights = tb.tsit5_dense_weights if solver == "tsit5" else tb.dopri5_dense_weights
    for _ in range(max_steps):
        if not tprev < t_end:
            break
        dt = tnext - tprev
        ks = [k1]
        for i in range(1, 7):
            Y = y + dt * sum(A[i][j] * ks[j] for j in range(i))
            if i == 6:
                y1 = Y
            ks.append(f(torch.tensor(tprev + C[i] * dt, dtype=torch.float64), Y))
        with torch.no_grad():
            err = dt * sum(E[i] * ks[i] for i in range(7))
            sc = atol + torch.maximum(y.abs(), y1.abs()) * rtol
            scaled = float(torch.sqrt(torch.mean((err / sc) ** 2)))
        keep = scaled < 1.0
        factor = 10.0 if scaled == 0 else 0.9 * scaled ** (-0.2)
        factor = min(max(factor, 1.0 if keep else 0.2), 10.0)
        dt_next = dt * factor
        if keep:
            while idx < T and float(ts[idx]) <= tnext:
                th = (float(ts[idx]) - tprev) / dt if dt != 0 else 0.0
                bw = dense_weights(th)
                out[idx] = y + dt * sum(bw[i] * ks[i] for i in range(7))
                idx += 1
            y, k1, tprev = y1, ks[6], tnext
            nacc += 1
        else:
            nrej += 1
        tprev = min(tprev, t_end)
        tn = tprev + dt_next
        if tn > t_end - 1e-10:
            tn = t_end if keep else tprev + 0.5 * (t_end - tprev)
        tnext = tn
    assert idx == T, "oracle: max_steps reached"
    return torch.stack(out), nacc, nrej


def neuralode_value_and_grad(weights, biases, activation, final_activation, max_time, ts, y0,
                             loss_fn, rtol=1e-3, atol=1e-6, dt0=None, stoich=None, solver="tsit5",
                             want_max_time=False):
    """weights / biases: lists of numpy arrays (eqx convention W [out,in]); ts [B,T]; y0 [B,S].
    loss_fn(ys [B,T,S] torch float64) -> scalar.  Returns (loss, ys, grads_W, grads_b, grad_y0,
    n_accepted [B], n_rejected [B])."""
    Ws = [torch.tensor(np.asarray(W, dtype=np.float64), requires_grad=True) for W in weights]
    bs = [None if b is None else torch.tensor(np.asarray(b, dtype=np.float64), requires_grad=True)
          for b in biases]
    y0t = torch.tensor(np.asarray(y0, dtype=np.float64), requires_grad=True)
    tst = torch.tensor(np.asarray(ts, dtype=np.float64))
    St = None if stoich is None else torch.tensor(np.asarray(stoich, dtype=np.float64), requires_grad=True)
    # (the reference's trainer makes max_time an array leaf of `func`, trainer.py:85-89: with
    #  want_max_time its gradient is appended to the returned tuple)
    mt = torch.tensor(float(max_time), dtype=torch.float64, requires_grad=True)
    f = mlp_field(Ws, bs, activation, final_activation, mt if want_max_time else float(max_time), St)
    ys, na, nr = [], [], []
    for b in range(y0t.shape[0]):
        step0 = float(tst[b, 1] - tst[b, 0]) if dt0 is None else dt0
        o, a_, r_ = solve_one(f, tst[b], y0t[b], rtol, atol, step0, solver=solver)
        ys.append(o); na.append(a_); nr.append(r_)
    ys = torch.stack(ys)
    loss = loss_fn(ys)
    nb = sum(b is not None for b in bs)
    params = Ws + [b for b in bs if b is not None] + [y0t] + ([St] if St is not None else []) \
        + ([mt] if want_max_time else [])
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    extra = ()
    if want_max_time:
        extra = (0.0 if grads[-1] is None else float(grads[-1]),)
        grads = grads[:-1]
    gW = [g.numpy() for g in grads[:len(Ws)]]
    gb_iter = iter(grads[len(Ws):len(Ws) + nb])
    gb = [None if b is None else next(gb_iter).numpy() for b in bs]
    gy0 = grads[len(Ws) + nb].numpy()
    if St is not None:
        return (float(loss.detach()), ys.detach().numpy(), gW, gb, gy0, np.array(na), np.array(nr),
                grads[-1].numpy()) + extra
    return (float(loss.detach()), ys.detach().numpy(), gW, gb, gy0, np.array(na), np.array(nr)) + extra


def uode_value_and_grad(weights, biases, activation, final_activation, max_time, ts, y0, loss_fn,
                        mech_states, mech_parameters, mech_exprs, theta_m, alpha_residual,
                        gate_matrix, rtol=1e-3, atol=1e-6, dt0=None):
    """UniversalODE (universalode.py:84-141): t0 = 0, field = mech + softplus(alpha) *
    sigmoid(10 G y) * MLP.  Returns (loss, ys, grads_W, grads_b, grad_theta_m, grad_alpha,
    grad_gate, n_accepted, n_rejected)."""
    Ws = [torch.tensor(np.asarray(W, dtype=np.float64), requires_grad=True) for W in weights]
    bs = [None if b is None else torch.tensor(np.asarray(b, dtype=np.float64), requires_grad=True)
          for b in biases]
    th = torch.tensor(np.asarray(theta_m, dtype=np.float64), requires_grad=True)
    al = torch.tensor(np.asarray(alpha_residual, dtype=np.float64), requires_grad=True)
    G = torch.tensor(np.asarray(gate_matrix, dtype=np.float64), requires_grad=True)
    y0t = torch.tensor(np.asarray(y0, dtype=np.float64))
    tst = torch.tensor(np.asarray(ts, dtype=np.float64))
    mech = sympy_field_torch(mech_states, mech_parameters, mech_exprs)
    ys, na, nr = [], [], []
    for b in range(y0t.shape[0]):
        f = mlp_field(Ws, bs, activation, final_activation, float(max_time),
                      uode=(mech, th, al, G, y0t[b]))
        step0 = float(tst[b, 1] - tst[b, 0]) if dt0 is None else dt0
        o, a_, r_ = solve_one(f, tst[b], y0t[b], rtol, atol, step0, max_steps=100000, t0=0.0)
        ys.append(o); na.append(a_); nr.append(r_)
    ys = torch.stack(ys)
    loss = loss_fn(ys)
    nb = sum(b is not None for b in bs)
    params = Ws + [b for b in bs if b is not None] + [th, al, G]
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    gW = [g.numpy() for g in grads[:len(Ws)]]
    gb_iter = iter(grads[len(Ws):len(Ws) + nb])
    gb = [None if b is None else next(gb_iter).numpy() for b in bs]
    z = lambda g, like: np.zeros_like(like.detach().numpy()) if g is None else g.numpy()
    return (float(loss.detach()), ys.detach().numpy(), gW, gb, z(grads[-3], th), z(grads[-2], al),
            z(grads[-1], G), np.array(na), np.array(nr))
