"""CUDA code generation for the neural vector fields (NeuralODE / RateFlowODE / UniversalODE).

Reference being replaced:
* ``catalax/neural/mlp.py:71-84``   ``MLP.__call__``: ``x = concat(y, [t / max_time])`` ->
  ``eqx.nn.MLP`` (``Linear(x) = W @ x + b``, ``W: [out, in]``; ``activation`` after every hidden
  layer, ``final_activation`` after the last; ``use_final_bias``), ``mlp.py:57-66``;
* ``catalax/neural/neuralode.py:49``        field = MLP;
* ``catalax/neural/rateflow.py:87-88``      field = ``stoich_matrix @ relu(MLP)`` (final relu :56);
* ``catalax/neural/universalode.py:84-101`` field = ``mech(t, y, theta) +
  softplus(alpha) * sigmoid(10 * G @ y) * MLP``.

The weights are one flat parameter array shared by the whole batch (``CTX_THETA_PTR``): the
integrator stages it in shared memory once per CTA and the generated ``ctx_rhs`` reads it with
128-bit broadcast loads; hidden activations live in registers (layers fully unrolled).  This is
the CUDA-core (FFMA) form of K3; the tcgen05 tile version is the planned replacement for
width >= 64.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import sympy as sp

from . import codegen as cg

ACTIVATIONS = {
    "tanh": "tanh({x})",
    "softplus": "(fmax({x}, R(0.0)) + log1p(exp(-fabs({x}))))",
    "celu": "(({x}) > R(0.0) ? ({x}) : expm1({x}))",
    "relu": "fmax({x}, R(0.0))",
    "identity": "({x})",
    "sigmoid": "(R(1.0) / (R(1.0) + exp(-({x}))))",
}


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


@dataclass
class MLPSpec:
    """Hyper-parameters of one neural vector field (names follow the reference constructors)."""

    kind: str                       # "neuralode" | "rateflow" | "universalode"
    data_size: int                  # S
    width_size: int
    depth: int                      # number of hidden layers
    out_size: int                   # S (neuralode/universalode) or reaction_size (rateflow)
    activation: str = "softplus"
    final_activation: str = "identity"
    use_final_bias: bool = False
    mech: Optional[cg.FieldSpec] = None      # universalode: mechanistic model (assignments inlined)

    def layer_dims(self) -> List[Tuple[int, int]]:
        dims = [self.data_size + 1] + [self.width_size] * self.depth + [self.out_size]
        return list(zip(dims[:-1], dims[1:]))


@dataclass
class MLPLayout:
    """Offsets (in elements) of everything packed into the flat parameter array."""

    w_off: List[int] = field(default_factory=list)
    b_off: List[int] = field(default_factory=list)
    in_pad: List[int] = field(default_factory=list)
    inv_max_time: int = 0
    stoich: int = -1
    gate: int = -1
    alpha: int = -1
    mech: int = -1
    size: int = 0


def layout_of(spec: MLPSpec) -> MLPLayout:
    lay = MLPLayout()
    off = 0
    for n_in, n_out in spec.layer_dims():
        ip = _pad4(n_in)
        lay.in_pad.append(ip)
        lay.w_off.append(off); off += n_out * ip
        lay.b_off.append(off); off += _pad4(n_out)
    lay.inv_max_time = off; off += 4
    S = spec.data_size
    if spec.kind == "rateflow":
        lay.stoich = off; off += _pad4(S * spec.out_size)
    if spec.kind == "universalode":
        lay.gate = off; off += _pad4(S * S)
        lay.alpha = off; off += _pad4(S)
        lay.mech = off; off += _pad4(len(spec.mech.parameters))
    lay.size = off
    return lay


def pack_parameters(spec: MLPSpec, weights: Sequence[np.ndarray], biases: Sequence[Optional[np.ndarray]],
                    max_time: float, dtype, stoich_matrix=None, gate_matrix=None, alpha_residual=None,
                    mech_parameters=None) -> np.ndarray:
    """Flatten eqx-convention weights (W [out, in], b [out]) into the kernel's layout."""
    lay = layout_of(spec)
    out = np.zeros(lay.size, dtype=np.float64)
    for l, ((n_in, n_out), W, b) in enumerate(zip(spec.layer_dims(), weights, biases)):
        W = np.asarray(W, dtype=np.float64)
        assert W.shape == (n_out, n_in), (W.shape, (n_out, n_in))
        blk = np.zeros((n_out, lay.in_pad[l]))
        blk[:, :n_in] = W
        out[lay.w_off[l]: lay.w_off[l] + blk.size] = blk.ravel()
        if b is not None:
            out[lay.b_off[l]: lay.b_off[l] + n_out] = np.asarray(b, dtype=np.float64)
    out[lay.inv_max_time] = 1.0 / float(max_time)
    S = spec.data_size
    if spec.kind == "rateflow":
        sm = np.asarray(stoich_matrix, dtype=np.float64)
        assert sm.shape == (S, spec.out_size)
        out[lay.stoich: lay.stoich + sm.size] = sm.ravel()
    if spec.kind == "universalode":
        out[lay.gate: lay.gate + S * S] = np.asarray(gate_matrix, dtype=np.float64).ravel()
        a = np.asarray(alpha_residual, dtype=np.float64)
        out[lay.alpha: lay.alpha + S] = np.logaddexp(a, 0.0)        # softplus(alpha_residual)
        mp = np.asarray(mech_parameters, dtype=np.float64)
        out[lay.mech: lay.mech + mp.size] = mp
    return out.astype(dtype)


def generate_mlp_field(spec: MLPSpec) -> cg.GeneratedField:
    S = spec.data_size
    lay = layout_of(spec)
    dims = spec.layer_dims()
    L = len(dims)
    src = [
        f"#define CTX_S {S}", f"#define CTX_P {lay.size}", "#define CTX_C 0",
        "#define CTX_THETA_PTR 1", f"#define CTX_P_PAD {lay.size}",
        "#if CTX_F64",
        "__device__ __forceinline__ void ctx_ld4(const real* p, real* o) {",
        "  const double2 a = *reinterpret_cast<const double2*>(p);",
        "  const double2 b = *reinterpret_cast<const double2*>(p + 2);",
        "  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; }",
        "#else",
        "__device__ __forceinline__ void ctx_ld4(const real* p, real* o) {",
        "  const float4 a = *reinterpret_cast<const float4*>(p);",
        "  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; }",
        "#endif",
        "__device__ __noinline__ void ctx_rhs(const real t, const real* y, const real* th,",
        "                                     const real* c, const real* y0, real* dy) {",
        "  (void)c; (void)y0;",
        f"  real h0[{lay.in_pad[0]}];",
    ]
    for i in range(S):
        src.append(f"  h0[{i}] = y[{i}];")
    src.append(f"  h0[{S}] = t * th[{lay.inv_max_time}];")
    for i in range(S + 1, lay.in_pad[0]):
        src.append(f"  h0[{i}] = R(0.0);")
    flops = 0
    JB = 8   # outputs per loop iteration: the loop body stays ~JB*(in + in/4) instructions
    for l, (n_in, n_out) in enumerate(dims):
        ip = lay.in_pad[l]
        last = l == L - 1
        act = spec.final_activation if last else spec.activation
        n_store = n_out if last else _pad4(n_out)
        # inputs are copied into registers (static indices in the unrolled k loop); outputs are
        # written with a run-time index, i.e. to a per-thread local array (L1-resident)
        src += [
            f"  real x{l}[{ip}];",
            "#pragma unroll",
            f"  for (int k = 0; k < {ip}; ++k) x{l}[k] = h{l}[k];",
            f"  real h{l + 1}[{n_store}];",
        ]
        body = [
            f"      real acc = th[{lay.b_off[l]} + j];",
            "#pragma unroll",
            f"      for (int k = 0; k < {ip}; k += 4) {{",
            f"        real w4[4]; ctx_ld4(th + {lay.w_off[l]} + j * {ip} + k, w4);",
            f"        acc += w4[0] * x{l}[k]; acc += w4[1] * x{l}[k + 1];",
            f"        acc += w4[2] * x{l}[k + 2]; acc += w4[3] * x{l}[k + 3];",
            "      }",
            f"      h{l + 1}[j] = " + ACTIVATIONS[act].format(x="acc") + ";",
        ]
        if n_out % JB == 0 and n_out > JB:
            src += ["#pragma unroll 1", f"  for (int j0 = 0; j0 < {n_out}; j0 += {JB}) {{",
                    "#pragma unroll", f"    for (int jj = 0; jj < {JB}; ++jj) {{",
                    "      const int j = j0 + jj;", *body, "    }", "  }"]
        else:
            src += ["#pragma unroll", f"  for (int j = 0; j < {n_out}; ++j) {{", "    {", *body, "    }", "  }"]
        for j in range(n_out, n_store):
            src.append(f"  h{l + 1}[{j}] = R(0.0);")
        flops += 2 * n_in * n_out + n_out
    o = f"h{L}"
    if spec.kind == "neuralode":
        for i in range(S):
            src.append(f"  dy[{i}] = {o}[{i}];")
    elif spec.kind == "rateflow":
        R_ = spec.out_size
        for i in range(S):
            terms = " + ".join(f"th[{lay.stoich + i * R_ + r}] * fmax({o}[{r}], R(0.0))" for r in range(R_))
            src.append(f"  dy[{i}] = {terms};")
        flops += 2 * S * R_
    elif spec.kind == "universalode":
        names = {cg.T_SYM: "t"}
        for i, s in enumerate(spec.mech.states):
            names[sp.Symbol(s)] = f"y[{i}]"
            names[sp.Symbol(f"{s}_init")] = f"y0[{i}]"
        for i, p in enumerate(spec.mech.parameters):
            names[sp.Symbol(p)] = f"th[{lay.mech + i}]"
        used = set().union(*[e.free_symbols for e in spec.mech.total_exprs()])
        if any(sp.Symbol(cn) in used for cn in spec.mech.constants):
            # the reference passes constants=None to the mechanistic stack (universalode.py:96)
            raise ValueError("UniversalODE: the mechanistic model must not use constants")
        lines, fm = cg._emit_block([(f"const real mech{i}", e) for i, e in
                                    enumerate(spec.mech.total_exprs())], names, "m")
        src += lines
        flops += fm
        for i in range(S):
            g = " + ".join(f"th[{lay.gate + i * S + j}] * y[{j}]" for j in range(S))
            src.append(f"  dy[{i}] = mech{i} + th[{lay.alpha + i}] * "
                       + ACTIVATIONS["sigmoid"].format(x=f"R(10.0) * ({g})") + f" * {o}[{i}];")
        flops += S * (2 * S + 8)
    else:
        raise ValueError(spec.kind)
    src += ["}", "#define CTX_D 0", "#define CTX_D_THETA 0", "#define CTX_D_Y0 0"]
    text = "\n".join(src) + "\n"
    key = hashlib.sha256(text.encode()).hexdigest()[:24]
    return cg.GeneratedField(text, S, lay.size, 0, 0, 0, 0, flops, 0, True, key)
