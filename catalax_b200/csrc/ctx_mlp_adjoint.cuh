// ctx_mlp_adjoint.cuh -- value-and-gradient of a NeuralODE solve w.r.t. the MLP weights.
//
// Replaces the reverse pass XLA builds for
//   catalax/neural/trainer.py:268-306   eqx.filter_value_and_grad(grad_loss): y_pred =
//                                        vmap(model)(ti, y0i), loss(y_pred, yi), d loss / d weights
//   catalax/neural/neuralode.py:38-59   diffeqsolve(ODETerm(MLP), Tsit5, t0=ts[0], t1=ts[-1], dt0,
//                                        y0, PIDController(rtol, atol), SaveAt(ts=ts))
// diffrax's default adjoint (RecursiveCheckpointAdjoint) differentiates the DISCRETE scheme: the
// accepted step sequence is fixed (the controller's factor is under stop_gradient), rejected
// steps and the error estimate do not enter, the dense-output polynomial does.  The same here:
//   ctx_mlp_adj_forward   integrates and records (t_n, dt_n, y_n, first save index, #saves) of
//                         every accepted step, and writes ys[B,T,S];
//   (host)                loss and its cotangent gys = d loss / d ys  (element-wise torch ops);
//   ctx_mlp_adj_backward  walks the recorded steps backwards: recomputes the 7 Tsit5 stages,
//                         pulls gys back through the interpolant, the step and the FSAL link, and
//                         through the MLP at every stage (vector-Jacobian product), accumulating
//                         d loss / d theta;
//   ctx_mlp_adj_reduce    sums the per-warp partial gradients in a fixed order (reproducible).
// Mapping: ONE WARP PER TRAJECTORY, lanes = hidden units (lane j owns units j, j + 32, ...).  The
// weights live in shared memory once per CTA, transposed and padded (Wt[i][j], row stride
// n_out + 1) so that both the forward product (lane = output unit) and the transposed product of
// the reverse pass (lane = input unit) read conflict-free; every warp owns a private gradient
// accumulator of the same shape (no atomics).  The trajectory state (S <= 8 values) is
// replicated in the registers of all lanes, so control flow is warp-uniform.
//
// Generated per model (tools/codegen_mlp.py): ADJ_L layers with adj_nin/adj_nout, offsets of the
// flat parameter layout (adj_woff, adj_boff, adj_inpad, ADJ_INVT), of the shared-memory pack
// (adj_soff: Wt block, then bias block), ADJ_SPACK, ADJ_WMAX, ADJ_ACT / ADJ_FACT, ADJ_P.
#pragma once

#if CTX_SOLVER == 0 || CTX_SOLVER == 1   // Tsit5 (the reference's default) and Dopri5 (trainer.py:128-132)

#ifndef ADJ_KIND
#define ADJ_KIND 0                           // 0 NeuralODE | 1 RateFlowODE
#endif
#define ADJ_UPL ((ADJ_WMAX + 31) / 32)       // units per lane
#define ADJ_NACT (ADJ_L + 1)                 // activation vectors of one MLP evaluation

struct ctx_adj_args {
  const real* y0;      // [B,S]
  const real* theta;   // [ADJ_P] flat parameter layout (codegen_mlp.layout_of)
  const real* ts;      // [B,T] or [T]
  real* ys;            // [B,T,S]            (forward: out, backward: unused)
  const real* gys;     // [B,T,S]            (backward: d loss / d ys)
  real* ck_real;       // [B,cap,2+S]        t_n, dt_n, y_n of every accepted step
  int* ck_int;         // [B,cap,2]          first save index, number of saves
  int* n_steps;        // [B]                accepted steps recorded
  int* status;         // [B]
  int* n_accepted;     // [B]
  int* n_rejected;     // [B]
  real* partial;       // [n_warps, ADJ_P]   per-warp gradient (backward)
  real* gy0;           // [B,S] or null      d loss / d y0 (backward)
  unsigned long long* work_counter;
  long long B;
  int T, cap, max_steps, stride_ts;
  real rtol, atol, dt0;
};

__device__ __forceinline__ real adj_act(const int id, const real x) {
  switch (id) {
    case 0: return tanh(x);
    case 1: return fmax(x, R(0.0)) + log1p(exp(-fabs(x)));
    case 2: return x > R(0.0) ? x : expm1(x);
    case 3: return fmax(x, R(0.0));
    default: return x;
  }
}
// derivative of the activation expressed through its OUTPUT h = act(z)
__device__ __forceinline__ real adj_dact(const int id, const real h) {
  switch (id) {
    case 0: return R(1.0) - h * h;                    // tanh
    case 1: return R(1.0) - exp(-h);                  // softplus: sigmoid(z) = 1 - exp(-h)
    case 2: return h > R(0.0) ? R(1.0) : h + R(1.0);  // celu (alpha = 1): exp(z) = h + 1
    case 3: return h > R(0.0) ? R(1.0) : R(0.0);      // relu
    default: return R(1.0);
  }
}

// shared-memory pack <- flat parameters (transposed, padded rows; biases behind every block)
__device__ __forceinline__ void adj_stage_weights(const real* th, real* pack) {
  for (int l = 0; l < ADJ_L; ++l) {
    const int ni = adj_nin[l], no = adj_nout[l];
    for (int e = threadIdx.x; e < ni * no; e += blockDim.x) {
      const int j = e / ni, i = e - j * ni;
      pack[adj_soff[l] + i * (no + 1) + j] = th[adj_woff[l] + j * adj_inpad[l] + i];
    }
    for (int j = threadIdx.x; j < no; j += blockDim.x)
      pack[adj_soff[l] + ni * (no + 1) + j] = th[adj_boff[l] + j];
  }
#if ADJ_KIND == 1
  for (int e = threadIdx.x; e < CTX_S * ADJ_R; e += blockDim.x) pack[ADJ_SSTOICH + e] = th[ADJ_STOICH + e];
#endif
#if ADJ_KIND == 2
  for (int e = threadIdx.x; e < CTX_S * CTX_S; e += blockDim.x) pack[ADJ_SGATE + e] = th[ADJ_GATE + e];
  for (int e = threadIdx.x; e < CTX_S; e += blockDim.x) pack[ADJ_SALPHA + e] = th[ADJ_ALPHA + e];
  for (int e = threadIdx.x; e < ADJ_PM; e += blockDim.x) pack[ADJ_SMECH + e] = th[ADJ_MECH + e];
#endif
}

// One MLP evaluation by a warp.  acts: [ADJ_NACT][ADJ_WMAX] in shared memory; acts[0] = (y, t/max_t),
// acts[l+1] = act_l(W_l acts[l] + b_l).  dy (replicated in all lanes) = acts[ADJ_L][0..S).
__device__ __forceinline__ void adj_mlp_forward(const real* pack, real* acts, const real t,
                                                const real* y, const real inv_max_t, real* dy,
                                                const int lane, const real* yinit) {
  (void)yinit;
  {
    real v = t * inv_max_t;
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) v = (i == lane) ? y[i] : v;
    if (lane <= CTX_S) acts[lane] = v;
  }
  __syncwarp();
#pragma unroll 1
  for (int l = 0; l < ADJ_L; ++l) {
    const int ni = adj_nin[l], no = adj_nout[l];
    const real* wt = pack + adj_soff[l];
    const real* hin = acts + l * ADJ_WMAX;
    real* hout = acts + (l + 1) * ADJ_WMAX;
    const int act = (l == ADJ_L - 1) ? ADJ_FACT : ADJ_ACT;
#pragma unroll
    for (int u = 0; u < ADJ_UPL; ++u) {
      const int j = lane + 32 * u;
      if (j < no) {
        real z = wt[ni * (no + 1) + j];
        for (int i = 0; i < ni; ++i) z += wt[i * (no + 1) + j] * hin[i];
        hout[j] = adj_act(act, z);
      }
    }
    __syncwarp();
  }
#if ADJ_KIND == 1   // RateFlowODE: dy = stoich [S,R] @ relu(MLP) (rateflow.py:87-88; the MLP's final
                    // activation is relu, the second relu is idempotent)
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) {
    real acc = R(0.0);
    for (int r = 0; r < ADJ_R; ++r) acc += pack[ADJ_SSTOICH + i * ADJ_R + r] * acts[ADJ_L * ADJ_WMAX + r];
    dy[i] = acc;
  }
#elif ADJ_KIND == 2   // UniversalODE (universalode.py:84-101):
                      // dy = mech(t, y, theta_m) + softplus(alpha) * sigmoid(10 G y) * MLP(t, y)
  {
    real m[CTX_S];
    adj_mech(t, y, pack + ADJ_SMECH, yinit, m);
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) {
      real z = R(0.0);
#pragma unroll
      for (int j = 0; j < CTX_S; ++j) z += pack[ADJ_SGATE + i * CTX_S + j] * y[j];
      const real g = R(1.0) / (R(1.0) + exp(-(R(10.0) * z)));
      dy[i] = m[i] + pack[ADJ_SALPHA + i] * g * acts[ADJ_L * ADJ_WMAX + i];
    }
  }
#else
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) dy[i] = acts[ADJ_L * ADJ_WMAX + i];
#endif
}

// Vector-Jacobian product of one MLP evaluation: fbar[S] (replicated) = cotangent of dy.
// Accumulates the weight / bias gradient into the warp's private pack `gp`; returns the cotangent
// of the input state in ybar[S] (replicated).  dl: [2][ADJ_WMAX] scratch in shared memory.
// g_invt accumulates the cotangent of 1 / max_time: the network's last input is t / max_time, and
// the reference's trainer turns max_time into a trainable array leaf (trainer.py:85-89).
__device__ __forceinline__ void adj_mlp_vjp(const real* pack, real* gp, const real* acts,
                                            real* dl, const real* fbar_in, real* ybar, const int lane,
                                            const real t, const real* yinit, real& g_invt) {
  (void)yinit;
#if ADJ_KIND == 2
  // gate / alpha / mechanistic part (replicated in all lanes; lane 0 accumulates the gradients):
  //   f = mech + a * g * o,  g = sigmoid(10 G y):  obar = fbar a g  (-> through the MLP below),
  //   abar = fbar g o,  zbar = 10 fbar a o g (1 - g),  Gbar = zbar (x) y,  ybar += G^T zbar + mech^T fbar
  real fbar[CTX_S], yex[CTX_S], Ys[CTX_S];
  {
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) { Ys[i] = acts[i]; yex[i] = R(0.0); }
    real thb[ADJ_PM > 0 ? ADJ_PM : 1];
#pragma unroll
    for (int q = 0; q < ADJ_PM; ++q) thb[q] = R(0.0);
    adj_mech_vjp(t, Ys, pack + ADJ_SMECH, yinit, fbar_in, yex, thb);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < ADJ_PM; ++q) gp[ADJ_SMECH + q] += thb[q];
    }
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) {
      real z = R(0.0);
#pragma unroll
      for (int j = 0; j < CTX_S; ++j) z += pack[ADJ_SGATE + i * CTX_S + j] * Ys[j];
      const real g = R(1.0) / (R(1.0) + exp(-(R(10.0) * z)));
      const real a_ = pack[ADJ_SALPHA + i];
      const real o_ = acts[ADJ_L * ADJ_WMAX + i];
      fbar[i] = fbar_in[i] * a_ * g;
      const real zb = R(10.0) * fbar_in[i] * a_ * o_ * g * (R(1.0) - g);
      if (lane == 0) {
        gp[ADJ_SALPHA + i] += fbar_in[i] * g * o_;
#pragma unroll
        for (int j = 0; j < CTX_S; ++j) gp[ADJ_SGATE + i * CTX_S + j] += zb * Ys[j];
      }
#pragma unroll
      for (int j = 0; j < CTX_S; ++j) yex[j] += pack[ADJ_SGATE + i * CTX_S + j] * zb;
    }
  }
#else
  const real* fbar = fbar_in;
#endif
  // delta' of the output layer
  {
    const int no = adj_nout[ADJ_L - 1];
#pragma unroll
    for (int u = 0; u < ADJ_UPL; ++u) {
      const int j = lane + 32 * u;
      if (j < no) {
        real f = R(0.0);
#if ADJ_KIND == 1
        // cotangent of the rates: stoich^T fbar; stoichiometry gradient: fbar (x) rates
        const real hj = acts[ADJ_L * ADJ_WMAX + j];
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) {
          f += pack[ADJ_SSTOICH + i * ADJ_R + j] * fbar[i];
          gp[ADJ_SSTOICH + i * ADJ_R + j] += fbar[i] * hj;
        }
#else
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) f = (j == i) ? fbar[i] : f;
#endif
        dl[j] = f * adj_dact(ADJ_FACT, acts[ADJ_L * ADJ_WMAX + j]);
      }
    }
    __syncwarp();
  }
  real* cur = dl;
  real* nxt = dl + ADJ_WMAX;
#pragma unroll 1
  for (int l = ADJ_L - 1; l >= 0; --l) {
    const int ni = adj_nin[l], no = adj_nout[l];
    const real* wt = pack + adj_soff[l];
    real* gw = gp + adj_soff[l];
    const real* hin = acts + l * ADJ_WMAX;
    // weight / bias gradient: lane = output unit j, private column of the transposed block
#pragma unroll
    for (int u = 0; u < ADJ_UPL; ++u) {
      const int j = lane + 32 * u;
      if (j < no) {
        const real d = cur[j];
        for (int i = 0; i < ni; ++i) gw[i * (no + 1) + j] += d * hin[i];
        gw[ni * (no + 1) + j] += d;
      }
    }
    // input cotangent: lane = input unit i
#pragma unroll
    for (int u = 0; u < ADJ_UPL; ++u) {
      const int i = lane + 32 * u;
      if (i < ni) {
        real s = R(0.0);
        for (int j = 0; j < no; ++j) s += wt[i * (no + 1) + j] * cur[j];
        nxt[i] = (l > 0) ? s * adj_dact(ADJ_ACT, hin[i]) : s;
      }
    }
    __syncwarp();
    real* tmp = cur; cur = nxt; nxt = tmp;
  }
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) ybar[i] = cur[i];
  g_invt += cur[CTX_S] * t;        // input CTX_S is t * (1 / max_time)
#if ADJ_KIND == 2
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) ybar[i] += yex[i];
#endif
  __syncwarp();
}

// Interpolation weights of the solver's dense output written as y + dt * sum_i b_i(theta) f_i.
#if CTX_SOLVER == 1
// Dopri5: diffrax evaluates Shampine's quartic through (y0, y1, ymid, f0, f1) (ctx_dense); with
// y1 = y0 + dt sum A7i f_i, ymid = y0 + dt sum M_i f_i, f0 = dt f_1, f1 = dt f_7 its y0 terms cancel
// and b_i(theta) = theta (d_i1 + theta (c_i + theta (b_i + theta a_i))),
//   a_i = 2 (d_i7 - d_i1) - 8 A7i + 16 M_i,  b_i = 5 d_i1 - 3 d_i7 + 14 A7i - 32 M_i,
//   c_i = d_i7 - 4 d_i1 - 5 A7i + 16 M_i                       (b_i(1) = A7i: the step itself).
__device__ __forceinline__ void adj_tsit5_weights(const real th, real* b) {
  const real A7[7] = {A71, A72, A73, A74, A75, A76, R(0.0)};
  const real M[7] = {M1, M2, M3, M4, M5, M6, M7};
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const real d1 = i == 0 ? R(1.0) : R(0.0), d7 = i == 6 ? R(1.0) : R(0.0);
    const real qa = R(2.0) * (d7 - d1) - R(8.0) * A7[i] + R(16.0) * M[i];
    const real qb = R(5.0) * d1 - R(3.0) * d7 + R(14.0) * A7[i] - R(32.0) * M[i];
    const real qc = d7 - R(4.0) * d1 - R(5.0) * A7[i] + R(16.0) * M[i];
    b[i] = th * (d1 + th * (qc + th * (qb + th * qa)));
  }
}
#else
// Tsit5 interpolation weights b_i(theta), i = 1..7 (same factored forms as ctx_dense)
__device__ __forceinline__ void adj_tsit5_weights(const real th, real* b) {
  const real t2 = th * th;
  b[0] = R(-1.0530884977290216) * th * (th - R(1.3299890189751412)) *
         (t2 - R(1.4364028541716351) * th + R(0.7139816917074209));
  b[1] = R(0.1017) * t2 * (t2 - R(2.1966568338249754) * th + R(1.2949852507374631));
  b[2] = R(2.490627285651252793) * t2 * (t2 - R(2.38535645472061657) * th + R(1.57803468208092486));
  b[3] = R(-16.54810288924490272) * (th - R(1.21712927295533244)) * (th - R(0.61620406037800089)) * t2;
  b[4] = R(47.37952196281928122) * (th - R(1.203071208372362603)) * (th - R(0.658047292653547382)) * t2;
  b[5] = R(-34.87065786149660974) * (th - R(1.2)) * (th - R(0.666666666666666667)) * t2;
  b[6] = R(2.5) * (th - R(1.0)) * (th - R(0.6)) * t2;
}
#endif

__device__ const real adj_A[7][6] = {
    {R(0.0), R(0.0), R(0.0), R(0.0), R(0.0), R(0.0)},
    {A21, R(0.0), R(0.0), R(0.0), R(0.0), R(0.0)},
    {A31, A32, R(0.0), R(0.0), R(0.0), R(0.0)},
    {A41, A42, A43, R(0.0), R(0.0), R(0.0)},
    {A51, A52, A53, A54, R(0.0), R(0.0)},
    {A61, A62, A63, A64, A65, R(0.0)},
    {A71, A72, A73, A74, A75, A76}};
__device__ const real adj_C[7] = {R(0.0), C2, C3, C4, C5, R(1.0), R(1.0)};

// stage states and slopes of one step from (t, dt, y): f[i] = MLP(t + c_i dt, Y_i); acts of stage
// i are kept at acts_all + i * act_stride for the reverse pass (stride 0: one scratch block, the
// forward kernel does not need them).  y1 = Y_7 = the step's result.
__device__ __forceinline__ void adj_stages(const real* pack, real* acts_all, const real t,
                                           const real dt, const real* y, const real inv_max_t,
                                           real (*f)[CTX_S], real* y1, const int lane,
                                           const bool have_f1, const int act_stride,
                                           const real* yinit) {
  real Y[CTX_S];
#pragma unroll
  for (int i = 0; i < 7; ++i) {   // (fully unrolled: static indices keep f in registers)
#pragma unroll
    for (int s = 0; s < CTX_S; ++s) {
      real acc = R(0.0);
#pragma unroll
      for (int j = 0; j < i; ++j) acc += adj_A[i][j] * f[j][s];
      Y[s] = y[s] + dt * acc;
    }
    if (i == 6) {
#pragma unroll
      for (int s = 0; s < CTX_S; ++s) y1[s] = Y[s];
    }
    if (i == 0 && have_f1) continue;   // FSAL: f[0] carried over from the previous step
    adj_mlp_forward(pack, acts_all + i * act_stride, t + adj_C[i] * dt, Y, inv_max_t, f[i], lane,
                    yinit);
  }
}

// ---------------------------------------------------------------------------- forward + record
extern "C" __global__ void __launch_bounds__(ADJ_WARPS * 32)
ctx_mlp_adj_forward(const ctx_adj_args a) {
  extern __shared__ __align__(16) unsigned char adj_smem[];
  real* pack = reinterpret_cast<real*>(adj_smem);
  real* acts = pack + ADJ_SPACK + (threadIdx.x >> 5) * (ADJ_NACT * ADJ_WMAX);
  adj_stage_weights(a.theta, pack);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const real inv_max_t = a.theta[ADJ_INVT];
  const int T = a.T;
  while (true) {
    unsigned long long bb = 0;
    if (lane == 0) bb = atomicAdd(a.work_counter, 1ull);
    bb = __shfl_sync(0xffffffffu, bb, 0);
    if (bb >= (unsigned long long)a.B) break;
    const long long b = (long long)bb;
    const real* ts = a.ts + (size_t)b * a.stride_ts;
    real* out = a.ys + (size_t)b * T * CTX_S;
    real* ckr = a.ck_real + (size_t)b * a.cap * (2 + CTX_S);
    int* cki = a.ck_int + (size_t)b * a.cap * 2;
    real y[CTX_S], ynew[CTX_S], f[7][CTX_S], yinit[CTX_S];
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) yinit[i] = y[i] = a.y0[(size_t)b * CTX_S + i];
    const real t_end = ts[T - 1];
#ifdef ADJ_T0_ZERO
    real tprev = R(0.0);                      // universalode.py:133: t0 = 0.0
#else
    real tprev = ts[0];                       // neuralode.py:52: t0 = ts[0]
#endif
    // dt0 <= 0: the reference's default dt0 = ts[1] - ts[0] of THIS trajectory (neuralode.py:44-47)
    const real dt0 = (a.dt0 > R(0.0) || T < 2) ? a.dt0 : ts[1] - ts[0];
    real tnext = fmin(tprev + dt0, t_end);
    int idx = 0, nacc = 0, nrej = 0, status = 0;
    if (!(tprev < t_end)) {                   // t0 == t1: every requested time is t0
      for (; idx < T; ++idx)
        if (lane < CTX_S) out[(size_t)idx * CTX_S + lane] = y[lane];
    }
    bool have_f1 = false;
    for (int n = 0; tprev < t_end; ++n) {
      if (n >= a.max_steps) { status = CTX_ST_MAX; break; }
      const real dt = tnext - tprev;
      adj_stages(pack, acts, tprev, dt, y, inv_max_t, f, ynew, lane, have_f1, 0, yinit);
      // error estimate / controller: same arithmetic as ctx_scaled_error + the I-controller
      bool has_nan = false;
#pragma unroll
      for (int i = 0; i < CTX_S; ++i) has_nan |= (ynew[i] != ynew[i]);
      real ssq = R(0.0);
#pragma unroll
      for (int i = 0; i < CTX_S; ++i) {
        real e = dt * (E1 * f[0][i] + E2 * f[1][i] + E3 * f[2][i] + E4 * f[3][i] + E5 * f[4][i] +
                       E6 * f[5][i] + E7 * f[6][i]);
        if (e != e) e = CTX_INF;
        const real y1 = has_nan ? y[i] : ynew[i];
        const real sc = a.atol + fmax(fabs(y[i]), fabs(y1)) * a.rtol;
        const real r = e / sc;
        ssq += r * r;
      }
      const real scaled = sqrt(ssq / (real)CTX_S);
      const bool keep = scaled < R(1.0);
      const real inv = R(1.0) / scaled;
      real factor = R(0.9) * ctx_pow02(inv);
      factor = fmin(fmax(factor, keep ? R(1.0) : R(0.2)), R(10.0));
      if (factor != factor) factor = inv;
      const real dt_next = dt * factor;
      if (keep) {
        const int idx0 = idx;
        while (idx < T) {
          const real tq = ts[idx];
          if (!(tq <= tnext)) break;
          const real den = tnext - tprev;
          const real th_ = (den == R(0.0)) ? R(0.0) : (tq - tprev) / den;
          real bw[7];
          adj_tsit5_weights(th_, bw);
          if (lane < CTX_S) {
            real v = R(0.0), yl = R(0.0);
#pragma unroll
            for (int i = 0; i < CTX_S; ++i) {
              if (i == lane) {
                yl = y[i];
                v = bw[0] * f[0][i] + bw[1] * f[1][i] + bw[2] * f[2][i] + bw[3] * f[3][i] +
                    bw[4] * f[4][i] + bw[5] * f[5][i] + bw[6] * f[6][i];
              }
            }
            out[(size_t)idx * CTX_S + lane] = yl + dt * v;
          }
          ++idx;
        }
        if (nacc < a.cap) {
          if (lane == 0) { ckr[nacc * (2 + CTX_S)] = tprev; ckr[nacc * (2 + CTX_S) + 1] = dt;
                           cki[nacc * 2] = idx0; cki[nacc * 2 + 1] = idx - idx0; }
          if (lane < CTX_S) {
            real yl = R(0.0);
#pragma unroll
            for (int i = 0; i < CTX_S; ++i) yl = (i == lane) ? y[i] : yl;
            ckr[nacc * (2 + CTX_S) + 2 + lane] = yl;
          }
        } else {
          status = 3;   // more accepted steps than checkpoint slots
        }
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) { y[i] = ynew[i]; f[0][i] = f[6][i]; }
        have_f1 = true;
        tprev = tnext;
        ++nacc;
      } else {
        ++nrej;   // f[0] = f(tprev, y) stays valid
        have_f1 = true;
      }
      tprev = fmin(tprev, t_end);
      real tn = tprev + dt_next;
      if (tn > t_end - CTX_TOL_END) tn = keep ? t_end : tprev + R(0.5) * (t_end - tprev);
      tnext = tn;
      if (status == 3) break;
      if (tprev < t_end && !(fabs(tnext) < CTX_INF)) { status = CTX_ST_BAD; break; }
    }
    if (tprev < t_end && status == 0) status = CTX_ST_MAX;
    for (; idx < T; ++idx)
      if (lane < CTX_S) out[(size_t)idx * CTX_S + lane] = CTX_INF;
    if (lane == 0) {
      a.status[b] = status;
      a.n_accepted[b] = nacc;
      a.n_rejected[b] = nrej;
      a.n_steps[b] = nacc < a.cap ? nacc : a.cap;
    }
  }
}

// ---------------------------------------------------------------------------- backward
extern "C" __global__ void __launch_bounds__(ADJ_WARPS * 32)
ctx_mlp_adj_backward(const ctx_adj_args a) {
  extern __shared__ __align__(16) unsigned char adj_smem[];
  real* pack = reinterpret_cast<real*>(adj_smem);
  const int wid = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  real* wbase = pack + ADJ_SPACK + wid * (ADJ_SPACK + 7 * ADJ_NACT * ADJ_WMAX + 2 * ADJ_WMAX);
  real* gp = wbase;                                   // private gradient pack
  real* acts = gp + ADJ_SPACK;                        // [7][ADJ_NACT][ADJ_WMAX]
  real* dl = acts + 7 * ADJ_NACT * ADJ_WMAX;          // [2][ADJ_WMAX]
  adj_stage_weights(a.theta, pack);
  real g_invt = R(0.0);      // d loss / d (1 / max_time) over this warp's trajectories (replicated)
  for (int e = lane; e < ADJ_SPACK; e += 32) gp[e] = R(0.0);
  __syncthreads();
  const real inv_max_t = a.theta[ADJ_INVT];
  const int T = a.T;
  // static assignment (warp w takes trajectories w, w + #warps, ...): the order in which a warp
  // accumulates its private gradient is fixed, hence the result is reproducible run to run
  const long long n_warps = (long long)gridDim.x * ADJ_WARPS;
  for (long long b = (long long)blockIdx.x * ADJ_WARPS + wid; b < a.B; b += n_warps) {
    const real* ts = a.ts + (size_t)b * a.stride_ts;
    const real* g = a.gys + (size_t)b * T * CTX_S;
    const real* ckr = a.ck_real + (size_t)b * a.cap * (2 + CTX_S);
    const int* cki = a.ck_int + (size_t)b * a.cap * 2;
    const int nsteps = a.n_steps[b];
    real lam[CTX_S], carry[CTX_S], yinit[CTX_S];
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) {
      lam[i] = R(0.0); carry[i] = R(0.0);
      yinit[i] = a.y0 ? a.y0[(size_t)b * CTX_S + i] : R(0.0);
    }
    if (a.status[b] == 0 && nsteps == 0) {
      // t0 == t1: ys[q] = y0 for every q
      for (int q = 0; q < T; ++q)
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) lam[i] += g[(size_t)q * CTX_S + i];
    }
    if (a.status[b] == 0)
    for (int n = nsteps - 1; n >= 0; --n) {
      const real t = ckr[n * (2 + CTX_S)], dt = ckr[n * (2 + CTX_S) + 1];
      const int idx0 = cki[n * 2], nsave = cki[n * 2 + 1];
      real y[CTX_S], y1[CTX_S], f[7][CTX_S], fb[7][CTX_S], yb[CTX_S];
#pragma unroll
      for (int i = 0; i < CTX_S; ++i) y[i] = ckr[n * (2 + CTX_S) + 2 + i];
      adj_stages(pack, acts, t, dt, y, inv_max_t, f, y1, lane, false, ADJ_NACT * ADJ_WMAX, yinit);
#pragma unroll
      for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) fb[k][i] = R(0.0);
#pragma unroll
      for (int i = 0; i < CTX_S; ++i) { fb[6][i] = carry[i]; yb[i] = R(0.0); }
      // dense output: ys[q] = y + dt * sum_k b_k(theta_q) f_k
      for (int q = idx0; q < idx0 + nsave; ++q) {
        const real th_ = (dt == R(0.0)) ? R(0.0) : (ts[q] - t) / dt;
        real bw[7];
        adj_tsit5_weights(th_, bw);
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) {
          const real gq = g[(size_t)q * CTX_S + i];
          yb[i] += gq;
#pragma unroll
          for (int k = 0; k < 7; ++k) fb[k][i] += dt * bw[k] * gq;
        }
      }
      // stage 7: f_7 = MLP(t + dt, y1), y1 = y + dt sum_j A7j f_j
      real Yb[CTX_S];
      adj_mlp_vjp(pack, gp, acts + 6 * (ADJ_NACT * ADJ_WMAX), dl, fb[6], Yb, lane, t + dt, yinit, g_invt);
#pragma unroll
      for (int i = 0; i < CTX_S; ++i) {
        const real tot = lam[i] + Yb[i];
        yb[i] += tot;
#pragma unroll
        for (int j = 0; j < 6; ++j) fb[j][i] += dt * adj_A[6][j] * tot;
      }
      // stages 6..2
#pragma unroll
      for (int k = 5; k >= 1; --k) {
        adj_mlp_vjp(pack, gp, acts + k * (ADJ_NACT * ADJ_WMAX), dl, fb[k], Yb, lane,
                    t + adj_C[k] * dt, yinit, g_invt);
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) {
          yb[i] += Yb[i];
#pragma unroll
          for (int j = 0; j < k; ++j) fb[j][i] += dt * adj_A[k][j] * Yb[i];
        }
      }
      // stage 1: f_1 is the previous step's f_7 (FSAL); the first step evaluates it itself
      if (n > 0) {
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) carry[i] = fb[0][i];
      } else {
        adj_mlp_vjp(pack, gp, acts, dl, fb[0], Yb, lane, t, yinit, g_invt);
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) yb[i] += Yb[i];
      }
#pragma unroll
      for (int i = 0; i < CTX_S; ++i) lam[i] = yb[i];
    }
    if (a.gy0 && lane < CTX_S) {
      real v = R(0.0);
#pragma unroll
      for (int i = 0; i < CTX_S; ++i) v = (i == lane) ? lam[i] : v;
      a.gy0[(size_t)b * CTX_S + lane] = v;
    }
  }
  // private pack -> flat parameter layout of this warp's partial row
  __syncwarp();
  real* prow = a.partial + ((size_t)blockIdx.x * ADJ_WARPS + wid) * ADJ_P;
  for (int e = lane; e < ADJ_P; e += 32) prow[e] = R(0.0);
  __syncwarp();
  if (lane == 0) prow[ADJ_INVT] = g_invt;
  for (int l = 0; l < ADJ_L; ++l) {
    const int ni = adj_nin[l], no = adj_nout[l];
    for (int e = lane; e < ni * no; e += 32) {
      const int j = e / ni, i = e - j * ni;
      prow[adj_woff[l] + j * adj_inpad[l] + i] = gp[adj_soff[l] + i * (no + 1) + j];
    }
    for (int j = lane; j < no; j += 32) prow[adj_boff[l] + j] = gp[adj_soff[l] + ni * (no + 1) + j];
  }
#if ADJ_KIND == 1
  for (int e = lane; e < CTX_S * ADJ_R; e += 32) prow[ADJ_STOICH + e] = gp[ADJ_SSTOICH + e];
#endif
#if ADJ_KIND == 2   // (d/d alpha is w.r.t. softplus(alpha_residual), the value the pack carries)
  for (int e = lane; e < CTX_S * CTX_S; e += 32) prow[ADJ_GATE + e] = gp[ADJ_SGATE + e];
  for (int e = lane; e < CTX_S; e += 32) prow[ADJ_ALPHA + e] = gp[ADJ_SALPHA + e];
  for (int e = lane; e < ADJ_PM; e += 32) prow[ADJ_MECH + e] = gp[ADJ_SMECH + e];
#endif
}

// grad[p] = sum over the per-warp rows (fixed order, double accumulation)
extern "C" __global__ void __launch_bounds__(128)
ctx_mlp_adj_reduce(const real* partial, real* grad, const int n_rows) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ADJ_P) return;
  double s = 0.0;
  for (int r = 0; r < n_rows; ++r) s += (double)partial[(size_t)r * ADJ_P + p];
  grad[p] = (real)s;
}

extern "C" __device__ const unsigned ctx_adj_info[5] = {
    (unsigned)ADJ_SPACK, (unsigned)(ADJ_NACT * ADJ_WMAX), (unsigned)ADJ_WARPS, (unsigned)ADJ_P,
    (unsigned)ADJ_WMAX};

#endif  // CTX_SOLVER == 0 || CTX_SOLVER == 1
