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
    rbf: bool = False                        # hidden activation = RBFLayer (mlp.py:86-118, rbf.py:32-39)

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
    rbf_off: List[int] = field(default_factory=list)   # per hidden layer: (g*W, 2g*sum mu, g*sum mu^2, 0)
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
    if spec.rbf:
        for _ in range(spec.depth):
            lay.rbf_off.append(off); off += 4
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
                    mech_parameters=None, rbf_mu=None, rbf_gamma=None) -> np.ndarray:
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
    if spec.rbf:
        # RBFLayer.__call__ (rbf.py:32-39) on a vector x [W] with centres mu [W]:
        #   out_i = exp(-gamma * sum_j (x_i - mu_j)^2) = exp(-(gamma W x_i^2 - 2 gamma sum(mu) x_i + gamma sum(mu^2)))
        for l in range(spec.depth):
            mu = np.asarray(rbf_mu[l], dtype=np.float64)
            g = float(rbf_gamma[l] if np.ndim(rbf_gamma) else rbf_gamma)
            out[lay.rbf_off[l]: lay.rbf_off[l] + 3] = (g * mu.size, 2.0 * g * mu.sum(), g * (mu * mu).sum())
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


ACT_IDS = {"tanh": 0, "softplus": 1, "celu": 2, "relu": 3, "identity": 4}


def tc_dims(spec: MLPSpec):
    """Padded dimensions of the tensor-core form, or None when the model does not qualify
    (N must be a multiple of 16 for M=128 tcgen05.mma, K a multiple of 8 for tf32)."""
    W = spec.width_size
    if W % 32 or W > 64 or spec.depth < 1 or spec.data_size + 1 > 8 or spec.out_size > 16:
        return None
    if spec.rbf:
        return None     # radial-basis hidden layers run the CUDA-core form
    in_pad, out_pad = 8, 16
    layers = [(in_pad if l == 0 else W, out_pad if l == spec.depth else W) for l in range(spec.depth + 1)]
    wpack = sum(2 * k * n for k, n in layers)
    aux_bias = sum(n for _, n in layers)
    extra = 4
    S = spec.data_size
    if spec.kind == "rateflow":
        extra += S * spec.out_size
    if spec.kind == "universalode":
        extra += S * S + S + len(spec.mech.parameters)
    aux = _pad4(aux_bias + extra)
    kmax = max(W, in_pad)
    smem = (wpack * 4 + 1023) // 1024 * 1024 + 4 * 128 * kmax * 4 + aux * 4 + 64
    tmem_cols = 32
    while tmem_cols < 2 * W:
        tmem_cols *= 2
    if smem > 227 * 1024:
        return None
    return dict(in_pad=in_pad, out_pad=out_pad, layers=layers, wpack=wpack, aux_bias=aux_bias,
                aux=aux, smem=smem, tmem_cols=tmem_cols)


def _umma_block(Wm: np.ndarray, n_pad: int, k_pad: int) -> np.ndarray:
    """[N,K] weights -> K-major / no-swizzle UMMA canonical layout: 8x4-float core matrices,
    consecutive along N (SBO = 128 B), k blocks n_pad/8 core matrices apart (LBO = n_pad*16 B)."""
    full = np.zeros((n_pad, k_pad), dtype=np.float32)
    full[: Wm.shape[0], : Wm.shape[1]] = Wm
    blk = full.reshape(n_pad // 8, 8, k_pad // 4, 4)         # [nb, r, kb, c]
    return np.ascontiguousarray(blk.transpose(2, 0, 1, 3)).ravel()   # [kb, nb, r, c]


def _tf32_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest (ties away, like cvt.rna.tf32.f32) to 10 explicit mantissa bits."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32)


def pack_tc_parameters(spec: MLPSpec, weights, biases, max_time, stoich_matrix=None,
                       gate_matrix=None, alpha_residual=None, mech_parameters=None) -> np.ndarray:
    """theta of ctx_solve_mlp_tc: [aux (biases, extras) | hi/lo TF32 weights in UMMA layout]."""
    d = tc_dims(spec)
    assert d is not None
    aux = np.zeros(d["aux"], dtype=np.float32)
    off = 0
    for (k, n), b in zip(d["layers"], biases):
        if b is not None:
            aux[off: off + len(b)] = np.asarray(b, dtype=np.float32)
        off += n
    ex = off
    aux[ex] = np.float32(1.0 / float(max_time))
    S = spec.data_size
    if spec.kind == "rateflow":
        aux[ex + 4: ex + 4 + S * spec.out_size] = np.asarray(stoich_matrix, dtype=np.float32).ravel()
    if spec.kind == "universalode":
        aux[ex + 4: ex + 4 + S * S] = np.asarray(gate_matrix, dtype=np.float32).ravel()
        aux[ex + 4 + S * S: ex + 4 + S * S + S] = np.logaddexp(np.asarray(alpha_residual, dtype=np.float64), 0.0)
        mp = np.asarray(mech_parameters, dtype=np.float32)
        aux[ex + 4 + S * S + S: ex + 4 + S * S + S + mp.size] = mp
    packs = []
    for (k, n), Wm in zip(d["layers"], weights):
        Wm = np.asarray(Wm, dtype=np.float32)
        hi = _tf32_round(Wm)
        lo = _tf32_round(Wm - hi)
        packs += [_umma_block(hi, n, k), _umma_block(lo, n, k)]
    return np.concatenate([aux] + packs).astype(np.float32)


def _tc_source(spec: MLPSpec) -> List[str]:
    d = tc_dims(spec)
    if d is None:
        return []
    S = spec.data_size
    ex = 4   # extras start after [inv_max_time, pad x3]
    out = [
        "#if !CTX_F64",
        "#define CTX_MLP_TC 1",
        f"#define MLP_S {S}", f"#define MLP_IN_PAD {d['in_pad']}", f"#define MLP_W {spec.width_size}",
        f"#define MLP_DEPTH {spec.depth}", f"#define MLP_OUT {spec.out_size}",
        f"#define MLP_OUT_PAD {d['out_pad']}", f"#define MLP_ACT {ACT_IDS[spec.activation]}",
        f"#define MLP_FINAL_ACT {ACT_IDS[spec.final_activation]}",
        f"#define MLP_AUX_SIZE {d['aux']}", f"#define MLP_AUX_EXTRA {d['aux_bias']}",
        f"#define MLP_WPACK_SIZE {d['wpack']}", f"#define TC_TMEM_COLS {d['tmem_cols']}",
        f"#define CTX_MLP_TC_SMEM {d['smem']}",
        "__device__ __forceinline__ void ctx_mlp_finish(const float t, const float* y, const float* y0,",
        "                                               const float* ex, const float* o, float* dy) {",
        "  (void)t; (void)y; (void)y0; (void)ex;",
    ]
    if spec.kind == "neuralode":
        out += [f"  dy[{i}] = o[{i}];" for i in range(S)]
    elif spec.kind == "rateflow":
        R_ = spec.out_size
        for i in range(S):
            out.append(f"  dy[{i}] = " + " + ".join(
                f"ex[{ex + i * R_ + r}] * fmaxf(o[{r}], 0.0f)" for r in range(R_)) + ";")
    else:
        names = {cg.T_SYM: "t"}
        for i, s in enumerate(spec.mech.states):
            names[sp.Symbol(s)] = f"y[{i}]"
            names[sp.Symbol(f"{s}_init")] = f"y0[{i}]"
        moff = ex + S * S + S
        for i, p in enumerate(spec.mech.parameters):
            names[sp.Symbol(p)] = f"ex[{moff + i}]"
        lines, _ = cg._emit_block([(f"const real mech{i}", e) for i, e in
                                   enumerate(spec.mech.total_exprs())], names, "q")
        out += lines
        for i in range(S):
            g = " + ".join(f"ex[{ex + i * S + j}] * y[{j}]" for j in range(S))
            out.append(f"  dy[{i}] = mech{i} + ex[{ex + S * S + i}] * "
                       f"(1.0f / (1.0f + expf(-(10.0f * ({g}))))) * o[{i}];")
    out += ["}", "#endif"]
    return out


def adjoint_dims(spec: MLPSpec):
    """Shared-memory pack of the weight-gradient kernels (csrc/ctx_mlp_adjoint.cuh), or None when
    the model is not a plain NeuralODE of at most 8 states (RateFlow / Universal fields: next)."""
    if spec.kind not in ("neuralode", "rateflow", "universalode") or spec.data_size > 8 or spec.rbf:
        return None
    if spec.kind != "rateflow" and spec.out_size != spec.data_size:
        return None
    dims = spec.layer_dims()
    wmax = max(max(ni, no) for ni, no in dims)
    wmax = (wmax + 31) // 32 * 32
    soff, off = [], 0
    for ni, no in dims:
        soff.append(off)
        off += ni * (no + 1) + no
    sstoich = off
    if spec.kind == "rateflow":
        off += spec.data_size * spec.out_size
    sgate = salpha = smech = off
    if spec.kind == "universalode":
        S_ = spec.data_size
        sgate = off; off += S_ * S_
        salpha = off; off += S_
        smech = off; off += len(spec.mech.parameters)
    nact = (len(dims) + 1) * wmax
    per_warp_bwd = off + 7 * nact + 2 * wmax

    def warps(word):
        w = (227 * 1024 // word - off) // per_warp_bwd
        return max(0, min(4, w))
    return dict(dims=dims, wmax=wmax, soff=soff, spack=off, sstoich=sstoich, sgate=sgate,
                salpha=salpha, smech=smech, nact=nact,
                per_warp_bwd=per_warp_bwd,
                warps32=warps(4), warps64=warps(8))


def _adjoint_mech_source(spec: MLPSpec, lay: MLPLayout, d) -> List[str]:
    """UniversalODE (universalode.py:84-101): the mechanistic field and its vector-Jacobian
    products w.r.t. the states and the mechanistic parameters, generated from the sympy model."""
    S = spec.data_size
    mech = spec.mech
    used = set().union(*[e.free_symbols for e in mech.total_exprs()])
    if any(sp.Symbol(cn) in used for cn in mech.constants):
        raise ValueError("UniversalODE: the mechanistic model must not use constants")
    names = {cg.T_SYM: "t"}
    y_syms, p_syms = [], []
    for i, s_ in enumerate(mech.states):
        names[sp.Symbol(s_)] = f"y[{i}]"
        names[sp.Symbol(f"{s_}_init")] = f"y0[{i}]"
        y_syms.append(sp.Symbol(s_))
    for i, p_ in enumerate(mech.parameters):
        names[sp.Symbol(p_)] = f"thm[{i}]"
        p_syms.append(sp.Symbol(p_))
    exprs = [sp.sympify(e) for e in mech.total_exprs()]
    fwd, _ = cg._emit_block([(f"m[{i}]", e) for i, e in enumerate(exprs)], names, "u")
    fb = [sp.Symbol(f"ctx_fb{i}") for i in range(S)]
    vnames = dict(names)
    for i, b in enumerate(fb):
        vnames[b] = f"fbar[{i}]"
    outs = []
    for j, ys in enumerate(y_syms):
        outs.append((f"const real yb{j}", sum(fb[i] * sp.diff(exprs[i], ys) for i in range(S))))
    for p_i, ps in enumerate(p_syms):
        outs.append((f"const real tb{p_i}", sum(fb[i] * sp.diff(exprs[i], ps) for i in range(S))))
    vjp, _ = cg._emit_block(outs, vnames, "v")
    P_m = len(mech.parameters)
    return [
        "#define ADJ_KIND 2", "#define ADJ_T0_ZERO 1", f"#define ADJ_PM {P_m}",
        f"#define ADJ_GATE {lay.gate}", f"#define ADJ_ALPHA {lay.alpha}", f"#define ADJ_MECH {lay.mech}",
        f"#define ADJ_SGATE {d['sgate']}", f"#define ADJ_SALPHA {d['salpha']}",
        f"#define ADJ_SMECH {d['smech']}",
        "__device__ __forceinline__ void adj_mech(const real t, const real* y, const real* thm,",
        "                                         const real* y0, real* m) {",
        "  (void)t; (void)y; (void)thm; (void)y0;", *fwd, "}",
        "__device__ __forceinline__ void adj_mech_vjp(const real t, const real* y, const real* thm,",
        "                                             const real* y0, const real* fbar, real* ybar,",
        "                                             real* thbar) {",
        "  (void)t; (void)y; (void)thm; (void)y0; (void)fbar; (void)thbar;", *vjp,
        *[f"  ybar[{j}] += yb{j};" for j in range(S)],
        *[f"  thbar[{p_i}] += tb{p_i};" for p_i in range(P_m)], "}",
    ]


def _adjoint_source(spec: MLPSpec) -> List[str]:
    d = adjoint_dims(spec)
    if d is None or d["warps32"] < 1:
        return []
    lay = layout_of(spec)
    arr = lambda name, vals: f"__device__ const int {name}[{len(vals)}] = {{" + ", ".join(map(str, vals)) + "};"
    out = [
        "#define CTX_MLP_ADJ 1",
        f"#define ADJ_L {len(d['dims'])}", f"#define ADJ_WMAX {d['wmax']}",
        f"#define ADJ_SPACK {d['spack']}", f"#define ADJ_P {lay.size}",
        f"#define ADJ_INVT {lay.inv_max_time}",
        f"#define ADJ_ACT {ACT_IDS[spec.activation]}", f"#define ADJ_FACT {ACT_IDS[spec.final_activation]}",
        "#if CTX_F64", f"#define ADJ_WARPS {max(1, d['warps64'])}", "#else",
        f"#define ADJ_WARPS {d['warps32']}", "#endif",
        arr("adj_nin", [ni for ni, _ in d["dims"]]), arr("adj_nout", [no for _, no in d["dims"]]),
        arr("adj_woff", lay.w_off), arr("adj_boff", lay.b_off), arr("adj_inpad", lay.in_pad),
        arr("adj_soff", d["soff"]),
    ]
    if spec.kind == "rateflow":
        out += ["#define ADJ_KIND 1", f"#define ADJ_R {spec.out_size}",
                f"#define ADJ_STOICH {lay.stoich}", f"#define ADJ_SSTOICH {d['sstoich']}"]
    if spec.kind == "universalode":
        out += _adjoint_mech_source(spec, lay, d)
    if d["warps64"] < 1:
        out = ["#if !CTX_F64"] + out + ["#endif"]
    return out


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
            f"      h{l + 1}[j] = " + (
                f"exp(-((th[{lay.rbf_off[l]}] * acc - th[{lay.rbf_off[l] + 1}]) * acc + th[{lay.rbf_off[l] + 2}]))"
                if (spec.rbf and not last) else ACTIVATIONS[act].format(x="acc")) + ";",
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
    src += _tc_source(spec)
    src += _adjoint_source(spec)
    text = "\n".join(src) + "\n"
    key = hashlib.sha256(text.encode()).hexdigest()[:24]
    return cg.GeneratedField(text, S, lay.size, 0, 0, 0, 0, flops, 0, True, key)
