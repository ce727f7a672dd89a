"""CPU restatement of the neural vector fields.  PARITY ORACLE (test infra).

Follows ``catalax/neural/mlp.py:71-84`` (``t/max_time``, ``concat(y, [t])``, eqx.nn.MLP =
``W @ x + b`` per layer, activation after hidden layers, final_activation after the last),
``rateflow.py:87-88`` (``stoich @ relu(mlp)``), ``universalode.py:84-101``
(``mech + softplus(alpha) * sigmoid(10 G y) * mlp``).  equinox (``pyproject.toml:26``) is not
vendored; ``eqx.nn.MLP``/``Linear`` conventions are restated from its documentation.
PARITY UNPINNED against equinox itself (the reference's neural tests are smoke tests).
"""
from __future__ import annotations

import numpy as np

ACT = {
    "tanh": np.tanh,
    "softplus": lambda x: np.logaddexp(x, 0.0),
    "celu": lambda x: np.maximum(x, 0.0) + np.expm1(np.minimum(x, 0.0)),
    "relu": lambda x: np.maximum(x, 0.0),
    "identity": lambda x: x,
}


def rbf_layer(x, mu, gamma):
    """``RBFLayer.__call__`` (catalax/neural/rbf.py:32-39) applied to each row of x [B, W] with
    centres mu [W]: ``diff = x[:, None] - mu[None, :]``, ``exp(-gamma * sum(diff**2, -1))`` --
    unit i sees the distances of ITS pre-activation to all W centres."""
    diff = x[:, :, None] - np.asarray(mu, dtype=x.dtype)[None, None, :]
    return np.exp(-x.dtype.type(gamma) * np.sum(np.square(diff), axis=-1))


def mlp_forward(x, weights, biases, activation, final_activation, rbf_mu=None, rbf_gamma=None):
    """x [B, in] -> [B, out]; W [out, in].  ``rbf_mu`` (one centre vector per hidden layer): the
    hidden activations are RBF layers and ``activation`` is the identity (mlp.py:86-118)."""
    h = x
    for l, (W, b) in enumerate(zip(weights, biases)):
        h = h @ W.T.astype(h.dtype)
        if b is not None:
            h = h + b.astype(h.dtype)
        last = l == len(weights) - 1
        if rbf_mu is not None and not last:
            h = rbf_layer(h, rbf_mu[l], rbf_gamma)
        else:
            h = ACT[final_activation if last else activation](h)
    return h


def make_rhs(kind, weights, biases, activation, final_activation, max_time, stoich_matrix=None,
             gate_matrix=None, alpha_residual=None, mech_rhs=None, mech_parameters=None,
             rbf_mu=None, rbf_gamma=None):
    def rhs(t, Y, theta, consts, y0):
        dt = Y.dtype
        x = np.concatenate([Y, (t / dt.type(max_time))[:, None]], axis=1)
        out = mlp_forward(x, weights, biases, activation, final_activation, rbf_mu, rbf_gamma)
        if kind == "neuralode":
            return out
        if kind == "rateflow":
            return np.maximum(out, 0.0) @ np.asarray(stoich_matrix, dtype=dt).T
        if kind == "universalode":
            B = Y.shape[0]
            mp = np.broadcast_to(np.asarray(mech_parameters, dtype=dt), (B, len(mech_parameters)))
            mech = mech_rhs(t, Y, mp, np.zeros((B, 0), dtype=dt), y0)
            gate = 1.0 / (1.0 + np.exp(-10.0 * (Y @ np.asarray(gate_matrix, dtype=dt).T)))
            alpha = np.logaddexp(np.asarray(alpha_residual, dtype=dt), 0.0)
            return mech + alpha[None, :] * gate * out
        raise ValueError(kind)
    return rhs
