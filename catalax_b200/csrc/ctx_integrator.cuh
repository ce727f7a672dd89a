// ctx_integrator.cuh -- batched adaptive-step ODE integrator template for sm_100a.
//
// Compiled at run time by NVRTC (see ctx_api.cu) AFTER the code-generated vector field of
// one model, which defines CTX_S/CTX_P/CTX_C/CTX_D and ctx_rhs()/ctx_rhs_tan().
//
// Replaces the XLA computation the reference triggers at
//   catalax/tools/simulation.py:259-273  diffeqsolve(ODETerm(stack), solver(), t0=0, t1=time[-1],
//                                         dt0, y0, args=(theta, consts, y0), SaveAt(ts=time),
//                                         PIDController(rtol, atol) | ConstantStepSize(),
//                                         max_steps, throw)
//   catalax/tools/simulation.py:346-355  jax.vmap(..., in_axes)
// Algorithm (diffrax 0.7, restated in oracle/diffrax_oracle.py): FSAL explicit RK step with
// increments k_i = f_i*dt, error = sum b_err_i k_i, I-controller
//   factor = clip(0.9 * scaled^(-1/5), keep ? 1 : 0.2, 10),  keep = scaled < 1,
// end-of-interval clipping (tol 1e-6 f32 / 1e-10 f64), dense output at ts on accepted steps,
// inf in every slot that was never reached.
//
// Compile-time switches (passed as -D by ctx_api.cu):
//   CTX_F64      0|1   arithmetic type (reference default is f32; enable_x64 -> f64)
//   CTX_SOLVER   0 tsit5 | 1 dopri5 | 2 euler | 3 heun      (simulation.py:23-26, simconfig.py:16)
//   CTX_TAN      0|1   carry the D forward tangents of ctx_rhs_tan through the same steps
#pragma once

#if CTX_F64
#define CTX_INF __longlong_as_double(0x7ff0000000000000LL)
#define CTX_QNAN __longlong_as_double(0x7ff8000000000000LL)
#define CTX_TOL_END 1e-10
#else
#define CTX_INF __int_as_float(0x7f800000)
#define CTX_QNAN __int_as_float(0x7fc00000)
#define CTX_TOL_END 1e-6f
#endif
#define CTX_ST_OK 0
#define CTX_ST_MAX 1
#define CTX_ST_BAD 2

#if CTX_TAN
#define CTX_N (CTX_S * (1 + CTX_D))
#else
#define CTX_N (CTX_S)
#endif

#if CTX_SOLVER == 0 || CTX_SOLVER == 1
#define CTX_NST 7
#define CTX_ADAPTIVE 1
#elif CTX_SOLVER == 2
#define CTX_NST 1
#define CTX_ADAPTIVE 0
#else
#define CTX_NST 2
#define CTX_ADAPTIVE 0
#endif

struct ctx_solve_args {
  const real* y0;      // [B,S] or [S]
  const real* theta;   // [B,P] or [P]
  const real* consts;  // [B,C] or [C]
  const real* ts;      // [B,T] or [T]
  real* ys;            // [B,T,S]
  real* sens;          // [B,T,S,D] or null        (jax.jacobian layout, simulation.py:315)
  int* status;         // [B]   0 ok | 1 max_steps reached | 2 non-finite step size
  int* n_accepted;     // [B]
  int* n_rejected;     // [B]
  unsigned long long* work_counter;  // persistent kernels: next unclaimed trajectory
  long long B;
  int T;
  int max_steps;
  int stride_y0, stride_theta, stride_consts, stride_ts;  // elements per row, 0 = shared
  int t0_from_ts;  // 0: integrate from t0 = 0 (simulation.py:262); 1: from ts[0] (neuralode.py:52)
  int inner;       // M > 0: b = d*M + m, theta row d, y0 / consts / ts row m (predictive.py:150-154)
  int q_div;       // persistent kernels: a warp claims clamp(remaining / q_div, 1, 32) trajectories
  int ts_smem;     // 1: the shared time grid (stride_ts == 0) is staged in shared memory
  real rtol, atol, dt0;   // dt0 <= 0: every trajectory takes ts[1] - ts[0] of ITS time row, the
                          // reference's default `dt0 or ts[1] - ts[0]` under vmap (neuralode.py:44-47)
  // CTX_LLSAVE builds (ctx_pointwise_loglik): ys receives log p(data | loc = y(t), |sigma|) instead
  // of the states -- loo.py:180-193 `likelihood(loc, scale).log_prob(data_safe)` per draw
  const real* ll_data;    // [M,T,S] observations (0 where masked: data_safe), row = the series
  const real* ll_sigma;   // [D,S]   noise scale of the draw (row = b / inner)
  int ll_likelihood;      // 0 Normal | 1 SoftLaplace
  // Scheduling hint of the persistent kernels: queue position u is trajectory order[u] (a
  // permutation of [0, B), usually "most expensive first" from the step counts of a previous solve
  // of similar inputs).  Results do not depend on it; null = index order.
  const int* order;
};
// first step size of a trajectory whose requested times are `row` (T of them)
// Divisions of the step loop are IEEE (what the reference's XLA path computes).  A/B switch for
// f32 builds: the SFU form (MUFU.RCP + FMUL, 2 ulp) instead of the IEEE sequence (RCP + 4 FFMA +
// FCHK + a slow-path CALL for zero / denormal / extreme operands): CTX_FAST_DIV bit 0 = the
// integrator's own divisions (error norm, 1/scaled, 1/dt), bit 1 = the divisions of the generated
// field (CTX_DIV in the generated source).  Measured on config 2 (B200, f32): T=16 1.34 -> 1.15 ms,
// T=1000 3.44 -> 3.30 ms, K2 1.57 -> 1.44 ms -- not taken: it changes f32 bits for 4 % of the
// headline.  f64 always divides in IEEE.
#ifndef CTX_FAST_DIV
#define CTX_FAST_DIV 0
#endif
#if !CTX_F64 && (CTX_FAST_DIV & 1)
#define CTX_IDIV(a, b) __fdividef((a), (b))
#else
#define CTX_IDIV(a, b) ((a) / (b))
#endif
#ifndef CTX_NORM_SKIP_CONST
#define CTX_NORM_SKIP_CONST 1
#endif
#ifndef CTX_CONST_MASK
#define CTX_CONST_MASK 0u     // bit i: the generated right-hand side of model state i is identically 0
#endif
__host__ __device__ constexpr bool ctx_rhs_is_zero(int i) { return i < 32 && ((CTX_CONST_MASK >> i) & 1u); }
// component i of the (augmented) state: model states first, then the tangents direction-major
// (comp = S * (1 + d) + state).  The tangent of a state with an identically-zero right-hand side
// has an identically-zero right-hand side too (its Jacobian row and parameter derivatives vanish).
#ifndef CTX_STAGE_SKIP_CONST
#define CTX_STAGE_SKIP_CONST 1
#endif
__host__ __device__ constexpr bool ctx_comp_is_const(int i) {
  return CTX_STAGE_SKIP_CONST && ctx_rhs_is_zero(i % CTX_S);
}
#define CTX_ROW_DT0(row, T) (((a.dt0 > R(0.0)) || (T) < 2) ? a.dt0 : ((row)[1] - (row)[0]))

// operand rows of trajectory b: `ro` for the parameters, `ri` for y0 / constants / times
#define CTX_ROWS(b, ro, ri)                                              \
  const long long ro = a.inner > 0 ? (b) / a.inner : (b);               \
  const long long ri = a.inner > 0 ? (b) - ro * a.inner : (b)

#ifndef CTX_THETA_PTR
#define CTX_THETA_PTR 0   // 1: parameters are a shared array too large for registers (MLP weights)
#endif

// ------------------------------------------------------------------ vector field
__device__ __forceinline__ void ctx_field(const real t, const real* Y, const real* th,
                                          const real* c, const real* y0, real* dY) {
#if CTX_TAN
  ctx_rhs_tan(t, Y, Y + CTX_S, th, c, y0, dY, dY + CTX_S);
#else
  ctx_rhs(t, Y, th, c, y0, dY);
#endif
}

// ------------------------------------------------------------------ tableaus
#if CTX_SOLVER == 0
// Tsitouras 2011 (diffrax Tsit5); b_err = b_sol - b_hat.
#define A21 R(0.161)
#define A31 R(-0.008480655492356988544426874250230774675121177393430391537369234245294192976164141156943)
#define A32 R(0.3354806554923569885444268742502307746751211773934303915373692342452941929761641411569)
#define A41 R(2.897153057105493432130432594192938764924887287701866490314866693455023795137503079289)
#define A42 R(-6.359448489975074843148159912383825625952700647415626703305928850207288721235210244366)
#define A43 R(4.362295432869581411017727318190886861027813359713760212991062156752264926097707165077)
#define A51 R(5.325864828439256604428877920840511317836476253097040101202360397727981648835607691791)
#define A52 R(-11.74888356406282787774717033978577296188744178259862899288666928009020615663593781589)
#define A53 R(7.495539342889836208304604784564358155658679161518186721010132816213648793440552049753)
#define A54 R(-0.09249506636175524925650207933207191611349983406029535244034750452930469056411389539635)
#define A61 R(5.861455442946420028659251486982647890394337666164814434818157239052507339770711679748)
#define A62 R(-12.92096931784710929170611868178335939541780751955743459166312250439928519268343184452)
#define A63 R(8.159367898576158643180400794539253485181918321135053305748355423955009222648673734986)
#define A64 R(-0.07158497328140099722453054252582973869127213147363544882721139659546372402303777878835)
#define A65 R(-0.02826905039406838290900305721271224146717633626879770007617876201276764571291579142206)
#define A71 R(0.09646076681806522951816731316512876333711995238157997181903319145764851595234062815396)
#define A72 R(0.01)
#define A73 R(0.4798896504144995747752495322905965199130404621990332488332634944254542060153074523509)
#define A74 R(1.379008574103741893192274821856872770756462643091360525934940067397245698027561293331)
#define A75 R(-3.290069515436080679901047585711363850115683290894936158531296799594813811049925401677)
#define A76 R(2.324710524099773982415355918398765796109060233222962411944060046314465391054716027841)
#define C2 R(0.161)
#define C3 R(0.327)
#define C4 R(0.9)
#define C5 R(0.9800255409045096857298102862870245954942137979563024768854764293221195950761080302604)
#define E1 R(0.0017800110522257773)
#define E2 R(0.0008164344596567463)
#define E3 R(-0.007880878010261994)
#define E4 R(0.1447110071732629)
#define E5 R(-0.5823571654525552)
#define E6 R(0.45808210592918686)
#define E7 R(-0.015151515151515152)
#elif CTX_SOLVER == 1
// Dormand-Prince 5(4) (diffrax Dopri5) + Shampine's mid-point weights.
#define A21 R(0.2)
#define A31 R(0.075)
#define A32 R(0.225)
#define A41 R(0.9777777777777777)
#define A42 R(-3.7333333333333334)
#define A43 R(3.5555555555555554)
#define A51 R(2.9525986892242035)
#define A52 R(-11.595793324188385)
#define A53 R(9.822892851699436)
#define A54 R(-0.2908093278463649)
#define A61 R(2.8462752525252526)
#define A62 R(-10.757575757575758)
#define A63 R(8.906422717743473)
#define A64 R(0.2784090909090909)
#define A65 R(-0.2735313036020583)
#define A71 R(0.09114583333333333)
#define A72 R(0.0)
#define A73 R(0.44923629829290207)
#define A74 R(0.6510416666666666)
#define A75 R(-0.322376179245283)
#define A76 R(0.13095238095238096)
#define C2 R(0.2)
#define C3 R(0.3)
#define C4 R(0.8)
#define C5 R(0.8888888888888888)
#define E1 R(0.0012326388888888873)
#define E2 R(0.0)
#define E3 R(-0.004252770290506136)
#define E4 R(0.036979166666666674)
#define E5 R(-0.05086379716981132)
#define E6 R(0.04190476190476192)
#define E7 R(-0.025)
#define M1 R(0.10013431883002395)
#define M2 R(0.0)
#define M3 R(0.3918321794184259)
#define M4 R(-0.02982460176594817)
#define M5 R(0.05893268337240795)
#define M6 R(-0.04497888809104361)
#define M7 R(0.023904308236133973)
#endif

// ------------------------------------------------------------------ one RK step
// K[i] holds f_i (not f_i*dt); the dt factor is applied once per stage sum.
// On entry K[0] = f(tprev, y).  On exit K[1..] are the stage slopes and ynew = y1.
__device__ __forceinline__ void ctx_rk_stages(const real tp, const real dt, const real* y,
                                              real (*K)[CTX_N], const real* th, const real* c,
                                              const real* y0, real* ynew) {
#if CTX_SOLVER == 0 || CTX_SOLVER == 1
  // (components with an identically-zero right-hand side keep their value: y + dt * 0)
  real yt[CTX_N];
#pragma unroll
  for (int i = 0; i < CTX_N; ++i) yt[i] = ctx_comp_is_const(i) ? y[i] : y[i] + dt * (A21 * K[0][i]);
  ctx_field(tp + C2 * dt, yt, th, c, y0, K[1]);
#pragma unroll
  for (int i = 0; i < CTX_N; ++i)
    yt[i] = ctx_comp_is_const(i) ? y[i] : y[i] + dt * (A31 * K[0][i] + A32 * K[1][i]);
  ctx_field(tp + C3 * dt, yt, th, c, y0, K[2]);
#pragma unroll
  for (int i = 0; i < CTX_N; ++i)
    yt[i] = ctx_comp_is_const(i) ? y[i] : y[i] + dt * (A41 * K[0][i] + A42 * K[1][i] + A43 * K[2][i]);
  ctx_field(tp + C4 * dt, yt, th, c, y0, K[3]);
#pragma unroll
  for (int i = 0; i < CTX_N; ++i)
    yt[i] = ctx_comp_is_const(i) ? y[i]
                                 : y[i] + dt * (A51 * K[0][i] + A52 * K[1][i] + A53 * K[2][i] + A54 * K[3][i]);
  ctx_field(tp + C5 * dt, yt, th, c, y0, K[4]);
#pragma unroll
  for (int i = 0; i < CTX_N; ++i)
    yt[i] = ctx_comp_is_const(i) ? y[i]
                                 : y[i] + dt * (A61 * K[0][i] + A62 * K[1][i] + A63 * K[2][i] +
                                                A64 * K[3][i] + A65 * K[4][i]);
  ctx_field(tp + dt, yt, th, c, y0, K[5]);
#pragma unroll
  for (int i = 0; i < CTX_N; ++i)
    ynew[i] = ctx_comp_is_const(i) ? y[i]
                                   : y[i] + dt * (A71 * K[0][i] + A72 * K[1][i] + A73 * K[2][i] +
                                                  A74 * K[3][i] + A75 * K[4][i] + A76 * K[5][i]);
  ctx_field(tp + dt, ynew, th, c, y0, K[6]);
#elif CTX_SOLVER == 2
#pragma unroll
  for (int i = 0; i < CTX_N; ++i) ynew[i] = y[i] + dt * K[0][i];
#else
  real yt[CTX_N];
#pragma unroll
  for (int i = 0; i < CTX_N; ++i) yt[i] = y[i] + dt * K[0][i];
  ctx_field(tp + dt, yt, th, c, y0, K[1]);
#pragma unroll
  for (int i = 0; i < CTX_N; ++i) ynew[i] = y[i] + dt * (R(0.5) * K[0][i] + R(0.5) * K[1][i]);
#endif
}

#if CTX_ADAPTIVE
// scaled^(-1/5) of the I-controller.  The factor only scales the next step size (it is clipped
// to [0.2 | 1, 10] and never enters the accept / reject decision), so f32 uses the SFU form
// exp2(0.2 * log2(x)) (~1e-7 relative, i.e. f32 rounding level; libdevice powf costs ~70
// instructions per step attempt); f64 keeps pow().  inf -> inf, 0 -> 0, NaN -> NaN as pow does.
__device__ __forceinline__ real ctx_pow02(const real x) {
#if CTX_F64
  return pow(x, 0.2);
#else
  return __powf(x, 0.2f);
#endif
}

// scaled error norm of PIDController: rms(y_err / (atol + rtol*max(|y0|,|y1|))) over the S
// model states only (tangents never enter the norm: the factor is under stop_gradient).
__device__ __forceinline__ real ctx_error_ssq(const real dt, const real* y, const real* ynew,
                                              const real (*K)[CTX_N], const real rtol,
                                              const real atol) {
  bool has_nan = false;
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) has_nan |= (ynew[i] != ynew[i]);
  real ssq = R(0.0);
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) {
#if CTX_NORM_SKIP_CONST
    // A state whose right-hand side is identically zero (CTX_CONST_MASK) has all-zero slopes:
    // its error estimate is dt * 0 and its term (0 / sc)^2 is exactly 0 -- except where that
    // quotient is NaN (sc zero or NaN, dt not finite), which is kept.  Skipping the division
    // leaves every bit as it was and keeps a zero numerator away from the slow path of the IEEE
    // division (FCHK -> CALL on every step attempt).
    if (ctx_rhs_is_zero(i)) {
      const real sc = atol + fabs(y[i]) * rtol;       // (ynew == y for such a state)
      const real nan_or_zero = dt - dt;               // NaN when dt is inf / NaN
      ssq += (sc == R(0.0) || sc != sc) ? CTX_QNAN : nan_or_zero;
      continue;
    }
#endif
    real e = dt * (E1 * K[0][i] + E2 * K[1][i] + E3 * K[2][i] + E4 * K[3][i] + E5 * K[4][i] +
                   E6 * K[5][i] + E7 * K[6][i]);
    if (e != e) e = CTX_INF;
    const real y1 = has_nan ? y[i] : ynew[i];
    const real sc = atol + fmax(fabs(y[i]), fabs(y1)) * rtol;
    const real r = CTX_IDIV(e, sc);
    ssq += r * r;
  }
  return ssq;
}
__device__ __forceinline__ real ctx_scaled_error(const real dt, const real* y, const real* ynew,
                                                 const real (*K)[CTX_N], const real rtol,
                                                 const real atol) {
  return sqrt(CTX_IDIV(ctx_error_ssq(dt, y, ynew, K, rtol, atol), (real)CTX_S));
}

// Accept / reject decision and raw step-size factor 0.9 * scaled^(-1/5) of one step attempt;
// `bad`: the error norm is NaN.  f64 (and CTX_FUSED_FACTOR=0): scaled = sqrt(ssq / S), keep =
// scaled < 1, factor through 1 / scaled and pow, as the reference writes it.  f32: the SAME
// decisions without the division, the square root and the reciprocal -- sqrt and the division by
// S are monotone with sqrt(1) = 1 and S / S = 1, so `scaled < 1` holds exactly when `ssq < S`; and
// the factor, which the SFU pow already computes only to f32 rounding level, is evaluated as
// 0.9 * 2^(-0.1 * (log2(ssq) - log2(S))).  0 -> inf, inf -> 0, NaN -> NaN as before.
#ifndef CTX_FUSED_FACTOR
#define CTX_FUSED_FACTOR 1
#endif
__device__ __forceinline__ real ctx_error_control(const real dt, const real* y, const real* ynew,
                                                  const real (*K)[CTX_N], const real rtol,
                                                  const real atol, bool& keep, bool& bad) {
  const real ssq = ctx_error_ssq(dt, y, ynew, K, rtol, atol);
#if !CTX_F64 && CTX_FUSED_FACTOR
  keep = ssq < (real)CTX_S;
  bad = ssq != ssq;
  return R(0.9) * exp2f(fmaf(-0.1f, __log2f(ssq), 0.1f * CTX_LOG2S));
#else
  const real scaled = sqrt(CTX_IDIV(ssq, (real)CTX_S));
  keep = scaled < R(1.0);
  bad = scaled != scaled;
  return R(0.9) * ctx_pow02(CTX_IDIV(R(1.0), scaled));
#endif
}
#endif

// Test hook (never defined in shipped builds): the step-size factor that follows loop iteration
// CTX_DBG_FACTOR_STEP is taken from constant CTX_DBG_FACTOR_CONST of the trajectory when that
// constant is positive.  In f32 the error estimate of the reference's first steps is rounding
// noise, so the factor the reference itself drew there is a nuisance parameter of every pin
// against reference OUTPUTS (tests/test_gpu_reference_pins.py fits it per trajectory, exactly as
// oracle.diffrax_oracle.solve(factor_override=...) does on the CPU).
#ifdef CTX_DBG_FACTOR_CONST
#define CTX_DBG_FACTOR(n_it, c, factor)                                                       \
  if ((n_it) == CTX_DBG_FACTOR_STEP && (c)[CTX_DBG_FACTOR_CONST] > R(0.0)) (factor) = (c)[CTX_DBG_FACTOR_CONST]
#else
#define CTX_DBG_FACTOR(n_it, c, factor)
#endif

// ------------------------------------------------------------------ dense output
// Evaluate the solver's interpolant of the accepted step [tp, tp+dt] at time tq.
__device__ __forceinline__ void ctx_dense(const real tq, const real tp, const real tn,
                                          const real* y, const real* ynew,
                                          const real (*K)[CTX_N], real* out) {
  const real den = tn - tp;
  const real dt = den;
  const real th = (den == R(0.0)) ? R(0.0) : CTX_IDIV(tq - tp, den);
#if CTX_SOLVER == 0
  const real t2 = th * th;
  const real b1 = R(-1.0530884977290216) * th * (th - R(1.3299890189751412)) *
                  (t2 - R(1.4364028541716351) * th + R(0.7139816917074209));
  const real b2 = R(0.1017) * t2 * (t2 - R(2.1966568338249754) * th + R(1.2949852507374631));
  const real b3 = R(2.490627285651252793) * t2 *
                  (t2 - R(2.38535645472061657) * th + R(1.57803468208092486));
  const real b4 = R(-16.54810288924490272) * (th - R(1.21712927295533244)) *
                  (th - R(0.61620406037800089)) * t2;
  const real b5 = R(47.37952196281928122) * (th - R(1.203071208372362603)) *
                  (th - R(0.658047292653547382)) * t2;
  const real b6 = R(-34.87065786149660974) * (th - R(1.2)) * (th - R(0.666666666666666667)) * t2;
  const real b7 = R(2.5) * (th - R(1.0)) * (th - R(0.6)) * t2;
#pragma unroll
  for (int i = 0; i < CTX_N; ++i)
    out[i] = y[i] + dt * (b1 * K[0][i] + b2 * K[1][i] + b3 * K[2][i] + b4 * K[3][i] +
                          b5 * K[4][i] + b6 * K[5][i] + b7 * K[6][i]);
#elif CTX_SOLVER == 1
#pragma unroll
  for (int i = 0; i < CTX_N; ++i) {
    const real ymid = y[i] + dt * (M1 * K[0][i] + M2 * K[1][i] + M3 * K[2][i] + M4 * K[3][i] +
                                   M5 * K[4][i] + M6 * K[5][i] + M7 * K[6][i]);
    const real f0 = dt * K[0][i], f1 = dt * K[6][i];
    const real y0_ = y[i], y1_ = ynew[i];
    const real a = R(2.0) * (f1 - f0) - R(8.0) * (y1_ + y0_) + R(16.0) * ymid;
    const real b = R(5.0) * f0 - R(3.0) * f1 + R(18.0) * y0_ + R(14.0) * y1_ - R(32.0) * ymid;
    const real c = f1 - R(4.0) * f0 - R(11.0) * y0_ - R(5.0) * y1_ + R(16.0) * ymid;
    out[i] = (((a * th + b) * th + c) * th + f0) * th + y0_;
  }
#else
#pragma unroll
  for (int i = 0; i < CTX_N; ++i) out[i] = y[i] + th * (ynew[i] - y[i]);
#endif
}

__device__ __forceinline__ void ctx_store_point(const ctx_solve_args& a, const long long b,
                                                const int idx, const real* v) {
  real* o = a.ys + ((size_t)b * a.T + idx) * CTX_S;
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) o[i] = v[i];
#if CTX_TAN
  if (a.sens) {
    real* s = a.sens + ((size_t)b * a.T + idx) * (CTX_S * CTX_D);
#pragma unroll
    for (int d = 0; d < CTX_D; ++d)
#pragma unroll
      for (int i = 0; i < CTX_S; ++i) s[i * CTX_D + d] = v[CTX_S + d * CTX_S + i];
  }
#endif
}

// ------------------------------------------------------------------ kernel v0
// One thread per trajectory, static assignment, direct stores.
#if !CTX_THETA_PTR || defined(CTX_NO_V1)   // weight-array models only get it when v1 cannot run
#define CTX_HAVE_V0 1
#endif
#ifdef CTX_HAVE_V0
extern "C" __global__ void __launch_bounds__(128) ctx_solve_v0(const ctx_solve_args a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;

  real y[CTX_N], ynew[CTX_N], y0[CTX_S];
  real c[CTX_C > 0 ? CTX_C : 1];
  real K[CTX_NST][CTX_N];
  CTX_ROWS(b, ro, ri);
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) y0[i] = y[i] = a.y0[(size_t)ri * a.stride_y0 + i];
#if CTX_THETA_PTR
  const real* th = a.theta + (size_t)ro * a.stride_theta;
#else
  real th[CTX_P > 0 ? CTX_P : 1];
#pragma unroll
  for (int i = 0; i < CTX_P; ++i) th[i] = a.theta[(size_t)ro * a.stride_theta + i];
#endif
#pragma unroll
  for (int i = 0; i < CTX_C; ++i) c[i] = a.consts[(size_t)ri * a.stride_consts + i];
#if CTX_TAN
#pragma unroll
  for (int i = CTX_S; i < CTX_N; ++i) y[i] = R(0.0);
#pragma unroll
  for (int i = 0; i < CTX_D_Y0; ++i) y[CTX_S + (CTX_D_THETA + i) * CTX_S + i] = R(1.0);
#endif
  const real* ts = a.ts + (size_t)ri * a.stride_ts;
  const int T = a.T;
  const real t_end = ts[T - 1];
  real tprev = a.t0_from_ts ? ts[0] : R(0.0);
  const real dt0_r = CTX_ROW_DT0(ts, T);
  real tnext = fmin(tprev + dt0_r, t_end);
  int idx = 0, nacc = 0, nrej = 0, status = 0;

  if (!(tprev < t_end)) {  // t0 == t1: every requested time is t0
    for (; idx < T; ++idx) ctx_store_point(a, b, idx, y);
  }
  ctx_field(tprev, y, th, c, y0, K[0]);

  for (int n = 0; tprev < t_end; ++n) {
    if (n >= a.max_steps) { status = 1; break; }
    const real dt = tnext - tprev;
    ctx_rk_stages(tprev, dt, y, K, th, c, y0, ynew);
    bool keep = true;
    real dt_next = dt0_r;
#if CTX_ADAPTIVE
    bool bad_norm = false;
    real factor = ctx_error_control(dt, y, ynew, K, a.rtol, a.atol, keep, bad_norm);
    factor = fmin(fmax(factor, keep ? R(1.0) : R(0.2)), R(10.0));
    if (bad_norm) factor = CTX_QNAN;  // clip(NaN) stays NaN in the reference
    CTX_DBG_FACTOR(n, c, factor);
    dt_next = dt * factor;
#endif
    if (keep) {
      real out[CTX_N];
      while (idx < T) {
        const real tq = ts[idx];
        if (!(tq <= tnext)) break;
        ctx_dense(tq, tprev, tnext, y, ynew, K, out);
        ctx_store_point(a, b, idx, out);
        ++idx;
      }
#pragma unroll
      for (int i = 0; i < CTX_N; ++i) y[i] = ynew[i];
#if CTX_SOLVER == 0 || CTX_SOLVER == 1
#pragma unroll
      for (int i = 0; i < CTX_N; ++i) K[0][i] = K[6][i];
#endif
      tprev = tnext;
      ++nacc;
    } else {
      ++nrej;
    }
#if !(CTX_SOLVER == 0 || CTX_SOLVER == 1)
    if (keep) ctx_field(tprev, y, th, c, y0, K[0]);
#endif
    tprev = fmin(tprev, t_end);
    real tn = tprev + dt_next;
    if (tn > t_end - CTX_TOL_END) tn = keep ? t_end : tprev + R(0.5) * (t_end - tprev);
    tnext = tn;
    if (tprev < t_end && !(fabs(tnext) < CTX_INF)) { status = 2; break; }
  }
  if (tprev < t_end && status == 0) status = 1;
  {
    real inf[CTX_N];
#pragma unroll
    for (int i = 0; i < CTX_N; ++i) inf[i] = i < CTX_S ? CTX_INF : R(0.0);
    for (; idx < T; ++idx) ctx_store_point(a, b, idx, inf);
  }
  a.status[b] = status;
  a.n_accepted[b] = nacc;
  a.n_rejected[b] = nrej;
}
#endif  // CTX_HAVE_V0

// =====================================================================================
// kernel v1: persistent warps, lane refill, warp-cooperative dense output
// =====================================================================================
// * Every lane integrates one trajectory.  When it finishes it claims the next unclaimed
//   trajectory from a global counter (one warp-aggregated atomicAdd per refill), so all 32
//   lanes of a warp execute a step attempt in every iteration regardless of how many
//   accepted / rejected steps their trajectories need ("step-count compaction").
// * Dense output is not evaluated by the lane that owns the step.  Each lane with an accepted
//   step that covers requested times publishes the step's interpolant as polynomial
//   coefficients in theta = (t - tprev)/dt in shared memory; the warp then cuts ALL pending save
//   points of ALL its lanes into groups of 4 consecutive points, deals the flattened groups out
//   to its 32 lanes (consecutive lanes = consecutive groups of one trajectory), evaluates the
//   polynomial (Horner) and writes each group with 128-bit stores: coalesced stores into
//   ys[B,T,S] although neighbouring lanes integrate unrelated trajectories.  (Kernels that also
//   write tangents use a point-wise variant staged through shared memory.)
// * inf-fill of failed trajectories and the t0 == t1 case reuse the same path (a constant
//   polynomial inf / y0).
//
// Interpolant as a polynomial.  Tsit5: sum_i b_i(theta) = theta, b_1'(0) = 1 and b_2..b_7 carry
// a factor theta^2, hence
//   y(theta) = y0 + dt*[theta*k_1 + sum_{i>=2} b_i(theta) (k_i - k_1)]
//            = y0 + theta*(a1 + theta*(a2 + theta*(a3 + theta*a4))),
//   a1 = dt*k_1,  a_m = dt * sum_{i=2..7} TB_im (k_i - k_1)   (TB_im: theta^m coefficient of b_i).
// Algebraically identical to diffrax's factored form; working on the differences k_i - k_1
// removes the cancellation between the O(10) weights, so it is at least as accurate in f32.
// Dopri5: diffrax's own quartic (a, b, c, f0, y0).  Euler/Heun: linear.
#ifndef CTX_WARPS
#define CTX_WARPS 4
#endif
#ifndef CTX_QUEUE
#define CTX_QUEUE 1   // 1: cp.async input queue | 0: lanes load their next trajectory directly
#endif
#ifndef CTX_MIN_BLOCKS
#define CTX_MIN_BLOCKS 1
#endif
#ifndef CTX_PASSES
#define CTX_PASSES 1      // save passes kept in flight together (see the save loop)
#endif
#ifndef CTX_GP
#define CTX_GP 4          // requested times per group (4 or 8): the unit of work of one lane in a pass
#endif
#define CTX_GSH (CTX_GP == 8 ? 3 : 2)
#ifndef CTX_ST256
#define CTX_ST256 1       // full-sector (256-bit) stores of complete groups, see the save loop
#endif
#ifndef CTX_DIRECT_SAVE
#define CTX_DIRECT_SAVE 0  // 1: the owning lane evaluates and stores its own points (sparse grids)
#endif
#ifndef CTX_BULK
// 1: complete groups leave through shared memory + one cp.async.bulk (TMA) per run of groups of
// one trajectory.  Measured on B200 (config 2, T=1000): 4.68 ms vs 4.44 ms (f32), 9.5 vs 8.1 ms
// (f64) for the direct 128-bit stores -- the runs are short (~480 B) and every copy is issued
// from uniform registers, one lane at a time -- so the direct stores stay the default.
#define CTX_BULK 0
#endif
#define CTX_FULL 0xffffffffu

#if CTX_SOLVER == 0 || CTX_SOLVER == 1
#define CTX_DEG 4
#else
#define CTX_DEG 1
#endif
#define CTX_NCOEF ((CTX_DEG + 1) * CTX_N)

// States with dy/dt == 0 identically (CTX_CONST_MASK, from the code generator) have the constant
// interpolant y: the step records of the non-tangent save path carry only the theta^0 coefficient
// for them ("packed" layout: [theta^0: all components][theta^m, m >= 1: dynamic components]).
#ifndef CTX_CONST_MASK
#define CTX_CONST_MASK 0u
#endif
#if CTX_TAN || CTX_S > 32
#define CTX_CMASK 0u
#else
#define CTX_CMASK (CTX_CONST_MASK)
#endif
__host__ __device__ constexpr int ctx_popc_c(unsigned m) { return m ? (int)(m & 1u) + ctx_popc_c(m >> 1) : 0; }
#define CTX_NDYN (CTX_N - ctx_popc_c(CTX_CMASK))
#define CTX_NPACK (CTX_N + CTX_DEG * CTX_NDYN)
__host__ __device__ constexpr bool ctx_is_const(int i) { return i < 32 && ((CTX_CMASK >> i) & 1u); }
// position of the theta^m coefficient of component i in the packed layout (i dynamic when m >= 1)
__host__ __device__ constexpr int ctx_pidx(int m, int i) {
  return m == 0 ? i : CTX_N + (m - 1) * CTX_NDYN + (i - ctx_popc_c(CTX_CMASK & ((1u << i) - 1u)));
}

#if CTX_SOLVER == 0
#define TB22 R(0.13169999999999999727)
#define TB23 R(-0.22339999999999999818)
#define TB24 R(0.1017)
#define TB32 R(3.9302962368947515285)
#define TB33 R(-5.9410338721315047347)
#define TB34 R(2.490627285651252793)
#define TB42 R(-12.411077166933676984)
#define TB43 R(30.338188630282321598)
#define TB44 R(-16.54810288924490272)
#define TB52 R(37.509313416511039195)
#define TB53 R(-88.178904894766401101)
#define TB54 R(47.37952196281928122)
#define TB62 R(-27.896526289197287806)
#define TB63 R(65.091894674793671526)
#define TB64 R(-34.87065786149660974)
#define TB72 R(1.5)
#define TB73 R(-4.0)
#define TB74 R(2.5)
#endif

// coef[m*CTX_N + i] = theta^m coefficient of component i on the accepted step [tp, tp+dt]
__device__ __forceinline__ void ctx_dense_coefs(const real dt, const real* y, const real* ynew,
                                                const real (*K)[CTX_N], real* coef) {
#pragma unroll
  for (int i = 0; i < CTX_N; ++i) {
#if CTX_SOLVER == 0
    const real d2 = K[1][i] - K[0][i], d3 = K[2][i] - K[0][i], d4 = K[3][i] - K[0][i],
               d5 = K[4][i] - K[0][i], d6 = K[5][i] - K[0][i], d7 = K[6][i] - K[0][i];
    coef[0 * CTX_N + i] = y[i];
    coef[1 * CTX_N + i] = dt * K[0][i];
    coef[2 * CTX_N + i] = dt * (TB22 * d2 + TB32 * d3 + TB42 * d4 + TB52 * d5 + TB62 * d6 + TB72 * d7);
    coef[3 * CTX_N + i] = dt * (TB23 * d2 + TB33 * d3 + TB43 * d4 + TB53 * d5 + TB63 * d6 + TB73 * d7);
    coef[4 * CTX_N + i] = dt * (TB24 * d2 + TB34 * d3 + TB44 * d4 + TB54 * d5 + TB64 * d6 + TB74 * d7);
#elif CTX_SOLVER == 1
    const real ymid = y[i] + dt * (M1 * K[0][i] + M2 * K[1][i] + M3 * K[2][i] + M4 * K[3][i] +
                                   M5 * K[4][i] + M6 * K[5][i] + M7 * K[6][i]);
    const real f0 = dt * K[0][i], f1 = dt * K[6][i];
    const real y0_ = y[i], y1_ = ynew[i];
    coef[0 * CTX_N + i] = y0_;
    coef[1 * CTX_N + i] = f0;
    coef[2 * CTX_N + i] = f1 - R(4.0) * f0 - R(11.0) * y0_ - R(5.0) * y1_ + R(16.0) * ymid;
    coef[3 * CTX_N + i] = R(5.0) * f0 - R(3.0) * f1 + R(18.0) * y0_ + R(14.0) * y1_ - R(32.0) * ymid;
    coef[4 * CTX_N + i] = R(2.0) * (f1 - f0) - R(8.0) * (y1_ + y0_) + R(16.0) * ymid;
#else
    coef[0 * CTX_N + i] = y[i];
    coef[1 * CTX_N + i] = ynew[i] - y[i];
#endif
  }
}

// the same coefficients in the packed layout (constant states: theta^0 only)
__device__ __forceinline__ void ctx_dense_coefs_packed(const real dt, const real* y, const real* ynew,
                                                       const real (*K)[CTX_N], real* pk) {
  real coef[CTX_NCOEF];
  ctx_dense_coefs(dt, y, ynew, K, coef);
#pragma unroll
  for (int i = 0; i < CTX_N; ++i) {
    pk[ctx_pidx(0, i)] = coef[i];
    if (!ctx_is_const(i)) {
#pragma unroll
      for (int m = 1; m <= CTX_DEG; ++m) pk[ctx_pidx(m, i)] = coef[m * CTX_N + i];
    }
  }
}

#ifndef CTX_LLSAVE
#define CTX_LLSAVE 0
#endif
#if CTX_LLSAVE
// numpyro's Normal / SoftLaplace log_prob (mcmc.py:59, 486-491) of x given loc and 1/|s|, log|s|
__device__ __forceinline__ real ctx_point_logp(const int likelihood, const real x, const real loc,
                                               const real inv_s, const real log_s) {
  const real z = (x - loc) * inv_s;
  if (likelihood == 0) return R(-0.5) * z * z - log_s - R(0.9189385332046727418);
  const real az = fabs(z);
  return R(-0.4515827052894548647) - log_s - (az + log1p(exp(R(-2.0) * az)));
}
#endif

#ifndef CTX_NO_V1
// Input queue of a persistent warp.  A finished lane must not wait ~1 us for the y0 / parameter
// rows of its next trajectory to arrive from HBM (the whole warp would wait with it, once per
// iteration: 32 lanes / ~30 steps per trajectory).  Every warp therefore claims trajectories in
// batches of q_batch, copies their operands with cp.async into one of two shared-memory buffers
// (lane = entry) and consumes one buffer while the other is in flight; the atomicAdd that claims
// the batch after next is issued a whole buffer ahead of its use as well.
#if CTX_THETA_PTR
#define CTX_QTH 0
#else
#define CTX_QTH CTX_P
#endif
#define CTX_QW (CTX_S + CTX_QTH + CTX_C + 2)   // y0, theta, constants, ts[0], ts[T-1]

__device__ __forceinline__ void ctx_cp_async(real* dst_smem, const real* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
#if CTX_F64
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
#endif
}
#define CTX_CP_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define CTX_CP_WAIT(n) asm volatile("cp.async.wait_group %0;" ::"n"(n) : "memory")

#if !CTX_TAN
// One published step ("segment"): everything a lane needs to evaluate 4 requested times of it,
// as one 16-byte-aligned record so that it moves with 128-bit shared-memory accesses.
struct alignas(16) ctx_rec {
  real v[2 + CTX_NPACK];   // tp, 1/dt, polynomial coefficients (packed layout, ctx_pidx)
  int i0, i1;              // [first, last) requested-time index covered by the step
  int gf;                  // (first group of the segment) - (its offset in the flat group list)
  int owner;               // publishing lane (index of the trajectory's pending-point buffer)
  unsigned row_lo, row_hi; // address of the trajectory's output row ys[b, 0, 0]
  int ri;                  // row of its time grid (per-row grids only)
};
#define CTX_REC_Q ((int)(sizeof(ctx_rec) / 16))
union ctx_rec_u {
  ctx_rec r;
  uint4 q[CTX_REC_Q];
  __device__ __forceinline__ ctx_rec_u() {}
};
#define CTX_PEND_Q ((int)(CTX_GP * CTX_S * sizeof(real) / 16))   // 16-byte chunks of one group
#endif

struct ctx_warp_smem {
  real q_in[2][CTX_QW][32];      // input queue: two buffers, one column per entry
  int q_idx[2][32];              // ... and the trajectory index of each entry (order hint applied)
#if !CTX_TAN
  ctx_rec rec[32];               // one record per publishing lane of this iteration
  uint4 pend[32][CTX_PEND_Q];    // per lane: the group of 4 points its trajectory has not completed
#if CTX_BULK
  uint4 stg[2 * CTX_PASSES][32][CTX_PEND_Q];   // staging of complete groups for the bulk stores
#endif
#else
  real rec[2 + CTX_NCOEF][32];   // tp, 1/dt, coefficients; one column per publishing lane
  long long ts_off[32];          // element offset into ts of the segment's save j = 0
  long long g_off[32];           // point index (b*T + idx) of the segment's save j = 0
  real stage_y[32 * CTX_S];      // interpolated states of one round, [point][state]
  real stage_s[32 * CTX_S * CTX_D];  // tangents of one round, [point][state][direction]
  long long gpt[32];             // element index (point*S) of each staged point (mixed rounds)
#endif
};

// Requested times.  SM: the grid is shared by all trajectories and staged in shared memory; it
// is then read through its 32-bit shared-space address `sa` (ld.shared: a generic pointer would
// cost a generic-to-shared conversion and the long-scoreboard path on every access).
template <bool SM>
__device__ __forceinline__ real ctx_ts_at(const real* row, const unsigned sa, const int i) {
  if (SM) {
    real v;
#if CTX_F64
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa + 8u * (unsigned)i));
#else
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sa + 4u * (unsigned)i));
#endif
    return v;
  }
  return row[i];
}

// First index in [lo, T) with ts[index] > t (ts ascending, ts[T-1] = t_end > t).  Requested
// times are usually an (almost) uniform grid: interpolate the position, fix it up with a few
// probes, and fall back to bisection of the remaining bracket for irregular grids.
template <bool SM>
__device__ __forceinline__ int ctx_upper_bound(const real* ts, const unsigned sa, int lo,
                                               const int T, const real t, const real t_end) {
  int hi = T - 1;                       // ts[hi] = t_end > t: the answer is in [lo, hi]
  const real t_lo = ctx_ts_at<SM>(ts, sa, lo);
  if (t_lo > t) return lo;
  // invariant from here: ts[lo] <= t < ts[hi]
  // (a position GUESS, verified by the probes below: the SFU division in f32 is enough for both
  //  dtypes and cannot change the result)
  const float frac = fminf(fmaxf(__fdividef((float)(t - t_lo), (float)(t_end - t_lo)), 0.0f), 1.0f);
  int g = lo + 1 + (int)(frac * (float)(hi - lo));
  g = min(max(g, lo + 1), hi);
#pragma unroll 1
  for (int probe = 0; probe < 3 && hi - lo > 1; ++probe) {
    if (ctx_ts_at<SM>(ts, sa, g) > t) { hi = g; g = max(g - 1, lo + 1); }
    else { lo = g; g = min(g + 1, hi); }
    if (hi - lo == 1) break;
  }
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ctx_ts_at<SM>(ts, sa, mid) > t) hi = mid; else lo = mid;
  }
  return hi;
}

// issue the operand copies of the batch [base, base + q_batch) into queue buffer `buf`
__device__ __forceinline__ void ctx_queue_fetch_q(const ctx_solve_args& a, real (*q_in)[CTX_QW][32],
                                                  int (*q_idx)[32], const int buf,
                                                  const unsigned long long base, const int count,
                                                  const int lane, const unsigned long long stride = 1ull) {
  const unsigned long long u = base + (unsigned long long)lane * stride;
  if (lane < count && u < (unsigned long long)a.B) {
    const long long bq = a.order ? (long long)a.order[u] : (long long)u;
    q_idx[buf][lane] = (int)bq;
    CTX_ROWS(bq, ro, ri);
    (void)ro;
    int k = 0;
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) ctx_cp_async(&q_in[buf][k++][lane], a.y0 + (size_t)ri * a.stride_y0 + i);
#if !CTX_THETA_PTR
#pragma unroll
    for (int i = 0; i < CTX_P; ++i) ctx_cp_async(&q_in[buf][k++][lane], a.theta + (size_t)ro * a.stride_theta + i);
#endif
#pragma unroll
    for (int i = 0; i < CTX_C; ++i) ctx_cp_async(&q_in[buf][k++][lane], a.consts + (size_t)ri * a.stride_consts + i);
    const real* tr = a.ts + (size_t)ri * a.stride_ts;
    ctx_cp_async(&q_in[buf][k++][lane], tr);
    ctx_cp_async(&q_in[buf][k++][lane], tr + (a.T - 1));
  }
  CTX_CP_COMMIT();
}

__device__ __forceinline__ void ctx_queue_fetch(const ctx_solve_args& a, ctx_warp_smem& w,
                                                const int buf, const unsigned long long base,
                                                const int count, const int lane,
                                                const unsigned long long stride = 1ull) {
  ctx_queue_fetch_q(a, w.q_in, w.q_idx, buf, base, count, lane, stride);
}

// TSM: the time grid is shared by all trajectories and staged in shared memory (ts_sm).
template <bool TSM>
__device__ __forceinline__ void ctx_solve_v1_body(const ctx_solve_args& a) {
  // dynamic smem: [CTX_THETA_PTR ? shared parameters (CTX_P_PAD reals)]
  //               [TSM ? time grid, padded to a multiple of 4 reals] [CTX_WARPS * ctx_warp_smem]
  extern __shared__ __align__(16) unsigned char ctx_dyn_smem[];
  unsigned char* sm_cur = ctx_dyn_smem;
#if CTX_THETA_PTR
  real* th_smem = reinterpret_cast<real*>(sm_cur);
  for (int i = threadIdx.x; i < CTX_P; i += blockDim.x) th_smem[i] = a.theta[i];
  const real* th = th_smem;
  sm_cur += CTX_P_PAD * sizeof(real);
#endif
  const int T = a.T;
  const real* ts_base = a.ts;     // base of all time rows (shared memory copy when TSM)
  unsigned ts_sa = 0;             // ... and its shared-space address
  if (TSM) {
    real* ts_sm = reinterpret_cast<real*>(sm_cur);
    ts_sa = (unsigned)__cvta_generic_to_shared(ts_sm);
    asm volatile("mov.u32 %0, %0;" : "+r"(ts_sa));   // opaque: keep it in a register
    const int Tpad = (T + CTX_GP - 1) & ~(CTX_GP - 1);
    for (int i = threadIdx.x; i < Tpad; i += blockDim.x) ts_sm[i] = a.ts[min(i, T - 1)];
    sm_cur += (size_t)Tpad * sizeof(real);
    ts_base = ts_sm;
  }
#if CTX_THETA_PTR
  __syncthreads();
#else
  if (TSM) __syncthreads();
#endif
  ctx_warp_smem& w = reinterpret_cast<ctx_warp_smem*>(sm_cur)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned le_mask = CTX_FULL >> (31 - lane);

  // per-lane trajectory state
  long long b = -1;           // trajectory index, -1 = lane is free
  int phase = 0;              // 0 integrating | 1 emit a fill segment then retire
  real y[CTX_N], ynew[CTX_N], y0[CTX_S];
#if !CTX_THETA_PTR
  real th[CTX_P > 0 ? CTX_P : 1];
#endif
  real c[CTX_C > 0 ? CTX_C : 1];
  real K[CTX_NST][CTX_N];
  real tprev = R(0.0), tnext = R(0.0), t_end = R(0.0);
  int idx = 0, nacc = 0, nrej = 0, status = 0;
  real dt0_l = a.dt0;         // first step size of this trajectory (the step size of Euler / Heun)
  const real* tsr = ts_base;
  long long ts_row = 0;       // element offset of this trajectory's time row
  int ri_row = 0;             // row of its y0 / constants / time grid
#if CTX_BULK && !CTX_TAN
  unsigned stg_round = 0;     // save rounds so far (selects the staging slots)
#endif

#if CTX_QUEUE
  // input queue (warp-uniform state).  Claim sizes follow the work that is left (guided
  // self-scheduling): whole warps of 32 while the batch is large, single trajectories at the
  // end, so that no warp sits on a private backlog while others have run dry.
  const unsigned long long Bq = (unsigned long long)a.B;
  const unsigned long long qdiv = (unsigned long long)a.q_div;
#define CTX_CLAIM(done) ((int)min(32ull, max(1ull, (Bq > (done) ? Bq - (done) : 0ull) / qdiv)))
  unsigned long long qb_cur = 0, qb_nxt = 0, qb_pend = 0;
  int qn_cur = CTX_CLAIM(0ull), qn_nxt = qn_cur, qn_pend = qn_cur;
  int q_cur = 0, q_pos = 0;
  bool q_empty = false;
  // With an order hint (most expensive rows first) the first two buffers of every warp are dealt
  // out STRIDED over the front of the queue -- warp g takes positions g, g + W, g + 2W, ... (W =
  // warps of the grid) -- so that the 32 W heaviest rows start at once, one per lane, instead of
  // the very heaviest all landing in the lanes of the first few warps; everything behind the
  // first 64 W positions is claimed from the shared counter in order, as without a hint.
  unsigned q_off = 0u;
  int qv_cur = 0, qv_nxt = 0;      // valid entries of the two buffers
  {
    if (a.order) {
      const unsigned Wq = gridDim.x * CTX_WARPS;
      q_off = 64u * Wq;
      qn_cur = qn_nxt = 32;
      qb_cur = (unsigned long long)blockIdx.x * CTX_WARPS + (threadIdx.x >> 5);
      qb_nxt = qb_cur + (unsigned long long)(32u * Wq);
    } else {
      if (lane == 0) qb_cur = atomicAdd(a.work_counter, (unsigned long long)(2 * qn_cur));
      qb_cur = __shfl_sync(CTX_FULL, qb_cur, 0);
      qb_nxt = qb_cur + qn_cur;
    }
    const unsigned long long qs0 = a.order ? (unsigned long long)(q_off / 64u) : 1ull;
    ctx_queue_fetch(a, w, 0, qb_cur, qn_cur, lane, qs0);
    ctx_queue_fetch(a, w, 1, qb_nxt, qn_nxt, lane, qs0);
    qv_cur = qb_cur < Bq ? (int)min((unsigned long long)qn_cur, (Bq - qb_cur + qs0 - 1ull) / qs0) : 0;
    qv_nxt = qb_nxt < Bq ? (int)min((unsigned long long)qn_nxt, (Bq - qb_nxt + qs0 - 1ull) / qs0) : 0;
    qn_pend = CTX_CLAIM(a.order ? (unsigned long long)q_off : qb_nxt + qn_nxt);
    if (lane == 0) qb_pend = atomicAdd(a.work_counter, (unsigned long long)qn_pend) + q_off;  // read one buffer later
    CTX_CP_WAIT(1);
    __syncwarp();
    q_empty = qb_cur >= Bq;
  }

  while (true) {
    // ------------------------------------------------------------- refill free lanes
    const unsigned need = __ballot_sync(CTX_FULL, b < 0);
    if (need && !q_empty) {
      const int nval = qv_cur;   // valid entries of this buffer
      const int take = min(__popc(need), nval - q_pos);
      const int rank = __popc(need & lt_mask);
      if (b < 0 && rank < take) {
        const int e = q_pos + rank;
        b = (long long)w.q_idx[q_cur][e];
        CTX_ROWS(b, ro, ri);
        (void)ro;
        ts_row = TSM ? 0 : ri * (long long)a.stride_ts;
        ri_row = (int)ri;
        int k = 0;
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) y0[i] = y[i] = w.q_in[q_cur][k++][e];
#if !CTX_THETA_PTR
#pragma unroll
        for (int i = 0; i < CTX_P; ++i) th[i] = w.q_in[q_cur][k++][e];
#endif
#pragma unroll
        for (int i = 0; i < CTX_C; ++i) c[i] = w.q_in[q_cur][k++][e];
        const real t_first = w.q_in[q_cur][k++][e];
        t_end = w.q_in[q_cur][k++][e];
#if CTX_TAN
#pragma unroll
        for (int i = CTX_S; i < CTX_N; ++i) y[i] = R(0.0);
#pragma unroll
        for (int i = 0; i < CTX_D_Y0; ++i) y[CTX_S + (CTX_D_THETA + i) * CTX_S + i] = R(1.0);
#endif
        tsr = ts_base + ts_row;
        tprev = a.t0_from_ts ? t_first : R(0.0);
        dt0_l = CTX_ROW_DT0(tsr, T);
        tnext = fmin(tprev + dt0_l, t_end);
        idx = 0; nacc = 0; nrej = 0; status = 0;
        phase = (tprev < t_end) ? 0 : 1;  // t0 == t1: emit y0 for every requested time
        ctx_field(tprev, y, th, c, y0, K[0]);
      }
      q_pos += take;
      if (q_pos >= nval) {
        // buffer consumed: the other one has been in flight for a whole buffer's worth of
        // trajectories; recycle this one for the batch claimed one rotation ago
        CTX_CP_WAIT(0);
        __syncwarp();
        const unsigned long long nbase = __shfl_sync(CTX_FULL, qb_pend, 0);
        ctx_queue_fetch(a, w, q_cur, nbase, qn_pend, lane);
        qb_cur = qb_nxt; qn_cur = qn_nxt; qv_cur = qv_nxt;
        qb_nxt = nbase; qn_nxt = qn_pend;
        qv_nxt = nbase < Bq ? (int)min((unsigned long long)qn_pend, Bq - nbase) : 0;
        qn_pend = CTX_CLAIM(nbase + qn_nxt);
        if (lane == 0 && nbase < Bq) qb_pend = atomicAdd(a.work_counter, (unsigned long long)qn_pend) + q_off;
        q_cur ^= 1; q_pos = 0;
        q_empty = qb_cur >= Bq;
      }
    }
#else   // direct refill: every lane loads its next trajectory from global memory itself
  bool q_empty = false;
  while (true) {
    const unsigned need = __ballot_sync(CTX_FULL, b < 0);
    if (need && !q_empty) {
      const int cnt = __popc(need);
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.work_counter, (unsigned long long)cnt);
      base = __shfl_sync(CTX_FULL, base, 0);
      if (base + cnt >= (unsigned long long)a.B) q_empty = true;
      if (b < 0) {
        const unsigned long long nb = base + __popc(need & lt_mask);
        if (nb < (unsigned long long)a.B) {
          b = a.order ? (long long)a.order[nb] : (long long)nb;
          CTX_ROWS(b, ro, ri);
          (void)ro;
          ts_row = TSM ? 0 : ri * (long long)a.stride_ts;
          ri_row = (int)ri;
#pragma unroll
          for (int i = 0; i < CTX_S; ++i) y0[i] = y[i] = a.y0[(size_t)ri * a.stride_y0 + i];
#if !CTX_THETA_PTR
#pragma unroll
          for (int i = 0; i < CTX_P; ++i) th[i] = a.theta[(size_t)ro * a.stride_theta + i];
#endif
#pragma unroll
          for (int i = 0; i < CTX_C; ++i) c[i] = a.consts[(size_t)ri * a.stride_consts + i];
#if CTX_TAN
#pragma unroll
          for (int i = CTX_S; i < CTX_N; ++i) y[i] = R(0.0);
#pragma unroll
          for (int i = 0; i < CTX_D_Y0; ++i) y[CTX_S + (CTX_D_THETA + i) * CTX_S + i] = R(1.0);
#endif
          tsr = ts_base + ts_row;
          t_end = tsr[T - 1];
          tprev = a.t0_from_ts ? tsr[0] : R(0.0);
          dt0_l = CTX_ROW_DT0(tsr, T);
          tnext = fmin(tprev + dt0_l, t_end);
          idx = 0; nacc = 0; nrej = 0; status = 0;
          phase = (tprev < t_end) ? 0 : 1;
          ctx_field(tprev, y, th, c, y0, K[0]);
        }
      }
    }
#endif  // CTX_QUEUE
    if (__ballot_sync(CTX_FULL, b >= 0) == 0u) {
      if (q_empty) break;
      continue;
    }

    // ------------------------------------------------------------- one step attempt
    const bool stepping = (b >= 0) && phase == 0;
    int nsave = 0;
    bool accepted = false, retire = false;
    real rec_tp = R(0.0), rec_dt = R(0.0);
    if (stepping) {
      const real dt = tnext - tprev;
      ctx_rk_stages(tprev, dt, y, K, th, c, y0, ynew);
      bool keep = true;
      real dt_next = dt0_l;
#if CTX_ADAPTIVE
      bool bad_norm = false;
      real factor = ctx_error_control(dt, y, ynew, K, a.rtol, a.atol, keep, bad_norm);
      factor = fmin(fmax(factor, keep ? R(1.0) : R(0.2)), R(10.0));
      CTX_DBG_FACTOR(nacc + nrej, c, factor);
      dt_next = dt * factor;
      if (bad_norm) dt_next = CTX_QNAN;  // NaN error norm poisons the step size
#endif
      rec_tp = tprev; rec_dt = dt;
      real tp_new = tprev;
      if (keep) {
        // number of requested times in (.., tnext]: first index in [idx, T) with ts > tnext
        if (!(tnext < t_end)) {
          nsave = T - idx;
        } else {
          nsave = ctx_upper_bound<TSM && !CTX_TAN>(tsr, ts_sa, idx, T, tnext, t_end) - idx;
        }
        tp_new = tnext;
        accepted = true;
        ++nacc;
      } else {
        ++nrej;
      }
      tp_new = fmin(tp_new, t_end);
      real tn = tp_new + dt_next;
      if (tn > t_end - CTX_TOL_END) tn = keep ? t_end : tp_new + R(0.5) * (t_end - tp_new);
      const bool done = !(tp_new < t_end);
      if (!done) {
        if (!(fabs(tn) < CTX_INF)) status = CTX_ST_BAD;
        else if (nacc + nrej >= a.max_steps) status = CTX_ST_MAX;
      }
      retire = done;
      tprev = tp_new; tnext = tn;
      phase = (status != 0) ? 1 : 0;  // failed: emit the inf fill in the next iteration
    }
    // a lane in phase 1 that did not step emits its fill segment now and retires
    const bool filling = (b >= 0) && !stepping;
    if (filling) { nsave = T - idx; retire = true; }

    // ------------------------------------------------------------- saves
#if CTX_DIRECT_SAVE && !CTX_TAN
    // Sparse output grids (a step covers a requested time only now and then: T <~ the number of
    // steps): the cooperative machinery below -- scan, records, passes, pending groups -- costs
    // ~40 % of the instructions of an iteration for a handful of points (ncu, T=16:
    // profiles/r2b_v1_T16.json).  Here the lane that owns the step evaluates its own points with
    // the same polynomial (same Horner order: bit-identical results) and stores them directly.
    if (nsave > 0) {
      real coef[CTX_NCOEF];
      real tp_ = R(0.0), inv_ = R(0.0);
      if (filling) {
        const bool fill_inf = status != 0;
#pragma unroll
        for (int i = 0; i < CTX_N; ++i) coef[i] = fill_inf ? CTX_INF : y[i];
#pragma unroll
        for (int i = CTX_N; i < CTX_NCOEF; ++i) coef[i] = R(0.0);
      } else {
        ctx_dense_coefs(rec_dt, y, ynew, K, coef);
        tp_ = rec_tp;
        inv_ = (rec_dt == R(0.0)) ? R(0.0) : CTX_IDIV(R(1.0), rec_dt);
      }
      real* o = a.ys + ((size_t)b * (size_t)T + (size_t)idx) * CTX_S;
#pragma unroll 1
      for (int j = 0; j < nsave; ++j) {
        const real tq = filling ? R(0.0) : (TSM ? ctx_ts_at<TSM>(tsr, ts_sa, idx + j) : tsr[idx + j]);
        const real th_ = (tq - tp_) * inv_;
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) {
          real v;
          if (ctx_is_const(i)) {
            v = coef[i];
          } else {
            v = coef[CTX_DEG * CTX_N + i];
#pragma unroll
            for (int mm = CTX_DEG - 1; mm >= 0; --mm) v = v * th_ + coef[mm * CTX_N + i];
          }
          o[(size_t)j * CTX_S + i] = v;
        }
      }
    }
#elif !CTX_TAN
    // Unified group path.  The requested times of every trajectory are cut into groups of 4
    // consecutive points aligned to multiples of 4; the groups touched by the accepted steps of
    // all lanes are flattened and dealt out to the 32 lanes (consecutive lanes = consecutive
    // groups of one segment).  A lane fetches its segment's record (6 x LDS.128 for 3 states),
    // evaluates the 4 points by Horner and writes a COMPLETE group with 128-bit stores.  A group
    // a step only partly covers is parked in the owner lane's pending buffer in shared memory
    // and merged by the step that completes it, so global memory only sees whole, aligned
    // 16-byte stores (scalar stores remain for unaligned rows and a ragged last group).
    const unsigned nz = __ballot_sync(CTX_FULL, nsave > 0);
    if (nz) {
      const int gq0 = idx >> CTX_GSH;
      const int ng = nsave > 0 ? ((idx + nsave + CTX_GP - 1) >> CTX_GSH) - gq0 : 0;
      int incl = ng;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(CTX_FULL, incl, d);
        if (lane >= d) incl += v;
      }
      const int goff = incl - ng;
      const int gtotal = __shfl_sync(CTX_FULL, incl, 31);
      if (nsave > 0) {
        const int s = __popc(nz & lt_mask);
        ctx_rec_u u;
        if (filling) {
          const bool fill_inf = status != 0;
#pragma unroll
          for (int i = 0; i < CTX_N; ++i) u.r.v[2 + ctx_pidx(0, i)] = fill_inf ? CTX_INF : y[i];
#pragma unroll
          for (int i = CTX_N; i < CTX_NPACK; ++i) u.r.v[2 + i] = R(0.0);
          u.r.v[0] = R(0.0);
          u.r.v[1] = R(0.0);
        } else {
          ctx_dense_coefs_packed(rec_dt, y, ynew, K, u.r.v + 2);
          u.r.v[0] = rec_tp;
          u.r.v[1] = (rec_dt == R(0.0)) ? R(0.0) : CTX_IDIV(R(1.0), rec_dt);
        }
        u.r.i0 = idx;
        u.r.i1 = idx + nsave;
        u.r.gf = gq0 - goff;
        u.r.owner = lane;
#ifdef CTX_DBG_ROWMASK    // developer experiment: all rows alias a small, L2-resident region
        const unsigned long long rowp = reinterpret_cast<unsigned long long>(
            a.ys + (size_t)(b & CTX_DBG_ROWMASK) * (size_t)T * CTX_S);
#else
        const unsigned long long rowp =
            reinterpret_cast<unsigned long long>(a.ys + (size_t)b * (size_t)T * CTX_S);
#endif
        u.r.row_lo = (unsigned)rowp;
        u.r.row_hi = (unsigned)(rowp >> 32);
#if CTX_LLSAVE
        u.r.ri = (int)b;      // the evaluating lane needs both rows: series = b % inner, draw = b / inner
#else
        u.r.ri = ri_row;
#endif
        uint4* dstq = reinterpret_cast<uint4*>(&w.rec[s]);
#pragma unroll
        for (int k = 0; k < CTX_REC_Q; ++k) dstq[k] = u.q[k];
      }
      __syncwarp();
      const unsigned myoff = nsave > 0 ? (unsigned)goff : 0xffffffffu;
      // One pass = 32 consecutive flat groups, one per lane.  Two ways to give a warp more
      // independent work per pass were measured on B200 (config 2, f32): CTX_PASSES=2 passes in
      // flight (144 registers, T=1000 -7 %, T=16 +15 %, f64 +25 %) and CTX_GP=8 requested times
      // per group (128 registers, T=1000 -11 %, T=250 +13 %, f64 +25 %).  The host compiles the
      // CTX_GP=8 variant for long f32 grids only (ctx_api.cu launch_solve).
#ifndef CTX_ST_OP
#define CTX_ST_OP ""      // cache operator of the 128-bit trajectory stores (".cs", ".cg", ...)
#endif
      struct pass_t {
        ctx_rec_u rr;
        real o[CTX_GP * CTX_S];
        int s_have, p0;
        bool act;
      } ps[CTX_PASSES];
#pragma unroll
      for (int k = 0; k < CTX_PASSES; ++k) {
        ps[k].s_have = -1;
        ps[k].rr.r.i0 = 0; ps[k].rr.r.i1 = 0; ps[k].rr.r.gf = 0; ps[k].rr.r.owner = 0;
        ps[k].rr.r.row_lo = 0; ps[k].rr.r.row_hi = 0; ps[k].rr.r.ri = 0;
      }
      int seg_base = 0;
      // bit j of bits[k]: a segment starts at flat group g0 + 32 k + j.  The masks of a round are
      // reduced one round ahead of their use, so that the REDUX latency is off the critical path.
      unsigned bits[CTX_PASSES];
#pragma unroll
      for (int k = 0; k < CTX_PASSES; ++k) {
        bits[k] = 0u;
        if (32 * k < gtotal) {
          const unsigned dn = myoff - (unsigned)(32 * k);
          bits[k] = __reduce_or_sync(CTX_FULL, dn < 32u ? (1u << dn) : 0u);
        }
      }
      for (int g0 = 0; g0 < gtotal; g0 += 32 * CTX_PASSES) {
        unsigned bits_cur[CTX_PASSES];
#pragma unroll
        for (int k = 0; k < CTX_PASSES; ++k) {
          bits_cur[k] = bits[k];
          if (g0 + 32 * (CTX_PASSES + k) < gtotal) {   // (warp-uniform)
            const unsigned dn = myoff - (unsigned)(g0 + 32 * (CTX_PASSES + k));
            bits[k] = __reduce_or_sync(CTX_FULL, dn < 32u ? (1u << dn) : 0u);
          }
        }
#pragma unroll
        for (int k = 0; k < CTX_PASSES; ++k) {
          pass_t& q = ps[k];
          ctx_rec_u& rr = q.rr;
          if (k > 0 && g0 + 32 * k >= gtotal) {   // (warp-uniform) no groups left for this pass
            q.act = false;
            continue;
          }
          const int s = seg_base + __popc(bits_cur[k] & le_mask) - 1;
          seg_base += __popc(bits_cur[k]);
          const int g = g0 + 32 * k + lane;
          const bool act = g < gtotal;
          q.act = act;
          if (act && s != q.s_have) {
            q.s_have = s;
            const uint4* srcq = reinterpret_cast<const uint4*>(&w.rec[s]);
#pragma unroll
            for (int j = 0; j < CTX_REC_Q; ++j) rr.q[j] = srcq[j];
          }
          const int p0 = (rr.r.gf + g) << CTX_GSH;
          q.p0 = p0;
          real* o = q.o;
          if (act) {
            real tq[CTX_GP];
            if (TSM) {   // padded to a multiple of CTX_GP and 16-byte aligned
#pragma unroll
              for (int h = 0; h < CTX_GP / 4; ++h) {
#if CTX_F64
                asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tq[4 * h]), "=d"(tq[4 * h + 1])
                    : "r"(ts_sa + 8u * (unsigned)p0 + 32u * h));
                asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tq[4 * h + 2]), "=d"(tq[4 * h + 3])
                    : "r"(ts_sa + 8u * (unsigned)p0 + 32u * h + 16u));
#else
                asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                    : "=f"(tq[4 * h]), "=f"(tq[4 * h + 1]), "=f"(tq[4 * h + 2]), "=f"(tq[4 * h + 3])
                    : "r"(ts_sa + 4u * (unsigned)p0 + 16u * h));
#endif
              }
            } else {
#if CTX_LLSAVE
              const int ri_t = a.inner > 0 ? rr.r.ri % a.inner : rr.r.ri;
#else
              const int ri_t = rr.r.ri;
#endif
              const real* r_ts = ts_base + (size_t)ri_t * a.stride_ts;
#pragma unroll
              for (int u = 0; u < CTX_GP; ++u) tq[u] = (p0 + u < rr.r.i1) ? r_ts[p0 + u] : rr.r.v[0];
            }
#pragma unroll
            for (int u = 0; u < CTX_GP; ++u) {
              const real th_ = (tq[u] - rr.r.v[0]) * rr.r.v[1];
#pragma unroll
              for (int i = 0; i < CTX_S; ++i) {
                if (ctx_is_const(i)) {
                  o[u * CTX_S + i] = rr.r.v[2 + ctx_pidx(0, i)];
                } else {
                  real v = rr.r.v[2 + ctx_pidx(CTX_DEG, i)];
#pragma unroll
                  for (int mm = CTX_DEG - 1; mm >= 0; --mm) v = v * th_ + rr.r.v[2 + ctx_pidx(mm, i)];
                  o[u * CTX_S + i] = v;
                }
              }
            }
#if CTX_LLSAVE
            {   // states -> point-wise log-likelihood of the observations at these times
              const int ro_l = a.inner > 0 ? rr.r.ri / a.inner : rr.r.ri;
              const int ri_l = a.inner > 0 ? rr.r.ri - ro_l * a.inner : rr.r.ri;
              real inv_s[CTX_S], log_s[CTX_S];
#pragma unroll
              for (int i = 0; i < CTX_S; ++i) {
                const real sa_ = fabs(a.ll_sigma[(size_t)ro_l * CTX_S + i]);
                inv_s[i] = R(1.0) / sa_;
                log_s[i] = log(sa_);
              }
              const real* drow = a.ll_data + ((size_t)ri_l * T + p0) * CTX_S;
#pragma unroll
              for (int u = 0; u < CTX_GP; ++u)
                if (p0 + u < T) {
#pragma unroll
                  for (int i = 0; i < CTX_S; ++i)
                    o[u * CTX_S + i] = ctx_point_logp(a.ll_likelihood, drow[u * CTX_S + i],
                                                      o[u * CTX_S + i], inv_s[i], log_s[i]);
                }
            }
#endif
            if (p0 < rr.r.i0) {   // points of this group that earlier steps evaluated
              union { uint4 q[CTX_PEND_Q]; real v[CTX_GP * CTX_S]; } pd;
#pragma unroll
              for (int j = 0; j < CTX_PEND_Q; ++j) pd.q[j] = w.pend[rr.r.owner][j];
#pragma unroll
              for (int u = 0; u < CTX_GP - 1; ++u)
                if (p0 + u < rr.r.i0) {
#pragma unroll
                  for (int i = 0; i < CTX_S; ++i) o[u * CTX_S + i] = pd.v[u * CTX_S + i];
                }
            }
          }
        }
#if CTX_BULK
        // The staging slots of this round were last read by the bulk copies committed two
        // rounds ago: at most one committed group (the previous round's) may still be pending.
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
#endif
        __syncwarp();   // every read of a pending buffer precedes the writes below
        bool ok[CTX_PASSES];
#pragma unroll
        for (int k = 0; k < CTX_PASSES; ++k) {
          pass_t& q = ps[k];
          const ctx_rec_u& rr = q.rr;
          const int p0 = q.p0;
          const real* o = q.o;
          ok[k] = false;
          if (q.act) {
            real* dst = reinterpret_cast<real*>(((unsigned long long)rr.r.row_hi << 32) |
                                                (unsigned long long)rr.r.row_lo) + p0 * CTX_S;
#ifdef CTX_DBG_NOSTORE   // developer experiment: everything but the global stores
            if (o[0] != R(-12345.678)) continue;
#endif
            if (p0 + CTX_GP <= rr.r.i1) {
#if CTX_ST256 && !CTX_BULK
              // Full-sector stores.  A group is 16*S (f32) / 32*S (f64) bytes; with 128-bit stores
              // and lanes 16*S bytes apart every store instruction writes HALF of 32 different
              // sectors, i.e. every sector travels from L1 to L2 twice -- and that path moves one
              // sector per clock per SM: measured, it (not DRAM, not the math) bounds the save
              // path.  sm_100 has 256-bit stores: f64 and even S write whole sectors directly; for
              // odd S the groups of even / odd index q start on / in the middle of a sector, so
              // lanes shift their 256-bit stores by one 16-byte piece according to the parity of
              // q, and the left-over pieces of two neighbouring lanes (last piece of the even
              // group, first of the odd one) share a sector AND an instruction: coalesced.
              if ((rr.r.row_lo & 31u) == 0u) {
#if CTX_F64
#pragma unroll
                for (int j = 0; j < CTX_GP * CTX_S / 4; ++j)
                  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j),
                               "d"(o[4 * j]), "d"(o[4 * j + 1]), "d"(o[4 * j + 2]), "d"(o[4 * j + 3])
                               : "memory");
#elif (CTX_GP * CTX_S) % 8 == 0
#pragma unroll
                for (int j = 0; j < CTX_GP * CTX_S / 8; ++j)
                  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * j),
                               "f"(o[8 * j]), "f"(o[8 * j + 1]), "f"(o[8 * j + 2]), "f"(o[8 * j + 3]),
                               "f"(o[8 * j + 4]), "f"(o[8 * j + 5]), "f"(o[8 * j + 6]), "f"(o[8 * j + 7])
                               : "memory");
#else
                const bool par = ((p0 >> 2) & 1) != 0;   // odd S, 4-point groups: starts mid-sector
#pragma unroll
                for (int j = 0; j < (CTX_S - 1) / 2; ++j) {
                  real v8[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) v8[e] = par ? o[8 * j + 4 + e] : o[8 * j + e];
                  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(
                                   dst + 8 * j + (par ? 4 : 0)),
                               "f"(v8[0]), "f"(v8[1]), "f"(v8[2]), "f"(v8[3]), "f"(v8[4]), "f"(v8[5]),
                               "f"(v8[6]), "f"(v8[7]) : "memory");
                }
                {
                  real v4[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) v4[e] = par ? o[e] : o[4 * (CTX_S - 1) + e];
                  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                                   dst + (par ? 0 : 4 * (CTX_S - 1))),
                               "f"(v4[0]), "f"(v4[1]), "f"(v4[2]), "f"(v4[3]) : "memory");
                }
#endif
              } else
#endif
              if ((rr.r.row_lo & 15u) == 0u) {   // p0 * S reals = a multiple of 16 bytes
#if CTX_BULK
                // stage the group; lanes are 4*S reals apart (conflict-free 128-bit stores)
                union { uint4 q[CTX_PEND_Q]; real v[CTX_GP * CTX_S]; } sg;
#pragma unroll
                for (int j = 0; j < CTX_GP * CTX_S; ++j) sg.v[j] = o[j];
#pragma unroll
                for (int j = 0; j < CTX_PEND_Q; ++j) w.stg[(stg_round & 1u) * CTX_PASSES + k][lane][j] = sg.q[j];
                ok[k] = true;
#elif CTX_F64
#pragma unroll
                for (int j = 0; j < CTX_GP * CTX_S / 2; ++j)
                  asm volatile("st.global" CTX_ST_OP ".v2.f64 [%0], {%1, %2};" ::"l"(dst + 2 * j), "d"(o[2 * j]),
                               "d"(o[2 * j + 1]) : "memory");
#else
#pragma unroll
                for (int j = 0; j < CTX_GP * CTX_S / 4; ++j)   // st.global: the address was rebuilt from integers
                  asm volatile("st.global" CTX_ST_OP ".v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j),
                               "f"(o[4 * j]), "f"(o[4 * j + 1]), "f"(o[4 * j + 2]), "f"(o[4 * j + 3])
                               : "memory");
#endif
              } else {
#pragma unroll
                for (int j = 0; j < CTX_GP * CTX_S; ++j) dst[j] = o[j];
              }
            } else if (rr.r.i1 == T) {   // ragged last group of the trajectory
#pragma unroll
              for (int u = 0; u < CTX_GP - 1; ++u)
                if (p0 + u < T) {
#pragma unroll
                  for (int i = 0; i < CTX_S; ++i) dst[u * CTX_S + i] = o[u * CTX_S + i];
                }
            } else {                      // incomplete: park it for the step that completes it
              union { uint4 q[CTX_PEND_Q]; real v[CTX_GP * CTX_S]; } pd;
#pragma unroll
              for (int j = 0; j < CTX_GP * CTX_S; ++j) pd.v[j] = o[j];
#pragma unroll
              for (int j = 0; j < CTX_PEND_Q; ++j) w.pend[rr.r.owner][j] = pd.q[j];
            }
          }
        }
#if CTX_BULK
        // One bulk copy (TMA) per run of staged groups of one trajectory: consecutive lanes =
        // consecutive groups = one contiguous, 16-byte aligned piece of ys[b].  The copy is
        // asynchronous (no store queue back-pressure on the warp) and L2 receives whole lines.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int k = 0; k < CTX_PASSES; ++k) {
          const pass_t& q = ps[k];
          const int s_prev = __shfl_up_sync(CTX_FULL, q.s_have, 1);
          const unsigned okm = __ballot_sync(CTX_FULL, ok[k]);
          const unsigned brk = __ballot_sync(CTX_FULL, lane == 0 || s_prev != q.s_have) | ~okm;
          if (ok[k] && (lane == 0 || s_prev != q.s_have || !((okm >> (lane - 1)) & 1u))) {
            const unsigned above = brk & ~(CTX_FULL >> (31 - lane));     // breaks at lanes > lane
            const int len = (above ? __ffs((int)above) - 1 : 32) - lane;
            real* dst = reinterpret_cast<real*>(((unsigned long long)q.rr.r.row_hi << 32) |
                                                (unsigned long long)q.rr.r.row_lo) + q.p0 * CTX_S;
            const unsigned src = (unsigned)__cvta_generic_to_shared(
                &w.stg[(stg_round & 1u) * CTX_PASSES + k][lane][0]);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                         "r"(src), "r"(len * (int)(CTX_GP * CTX_S * sizeof(real))) : "memory");
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        ++stg_round;
#endif
      }
    }
#else   // CTX_TAN: batched point path (states + tangents staged through shared memory)
    const unsigned nz = __ballot_sync(CTX_FULL, nsave > 0);
    if (nz) {
      int incl = nsave;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(CTX_FULL, incl, d);
        if (lane >= d) incl += v;
      }
      const int off = incl - nsave;
      const int total = __shfl_sync(CTX_FULL, incl, 31);
      if (nsave > 0) {
        const int s = __popc(nz & lt_mask);
        real coef[CTX_NCOEF];
        if (filling) {
          const bool fill_inf = status != 0;
#pragma unroll
          for (int i = 0; i < CTX_N; ++i)
            coef[i] = fill_inf ? (i < CTX_S ? CTX_INF : R(0.0)) : y[i];
#pragma unroll
          for (int i = CTX_N; i < CTX_NCOEF; ++i) coef[i] = R(0.0);
          w.rec[0][s] = R(0.0);
          w.rec[1][s] = R(0.0);
        } else {
          ctx_dense_coefs(rec_dt, y, ynew, K, coef);
          w.rec[0][s] = rec_tp;
          w.rec[1][s] = (rec_dt == R(0.0)) ? R(0.0) : CTX_IDIV(R(1.0), rec_dt);
        }
#pragma unroll
        for (int i = 0; i < CTX_NCOEF; ++i) w.rec[2 + i][s] = coef[i];
        w.ts_off[s] = ts_row + idx - off;
        w.g_off[s] = b * (long long)T + idx - off;
      }
      __syncwarp();
      if (total > 0) {
      // myoff: where this lane's segment starts in the flattened list (none: never matches)
      const unsigned myoff = nsave > 0 ? (unsigned)off : 0xffffffffu;
      int seg_base = 0, s_have = -1;
      real r_tp = R(0.0), r_inv = R(0.0), r_coef[CTX_NCOEF];
      const real* tq_ptr = ts_base;    // &ts[segment offset + j] of this lane's point
      long long r_ebase = 0;           // (g_off[s]) * S: element index of the segment's j = 0
      for (int r0 = 0; r0 < total; r0 += 32) {
        const unsigned dpos = myoff - (unsigned)r0;
        const unsigned bits = __reduce_or_sync(CTX_FULL, dpos < 32u ? (1u << dpos) : 0u);
        const int s = seg_base + __popc(bits & le_mask) - 1;
        seg_base += __popc(bits);
        const int j = r0 + lane;
        const int cnt = min(32, total - r0);
        const bool uniform = (bits & ~1u) == 0u;  // every point of this round: same segment
        if (s != s_have) {  // (lanes past `total` follow the last segment: harmless)
          s_have = s;
          r_tp = w.rec[0][s]; r_inv = w.rec[1][s];
#pragma unroll
          for (int i = 0; i < CTX_NCOEF; ++i) r_coef[i] = w.rec[2 + i][s];
          tq_ptr = ts_base + (w.ts_off[s] + j);
          r_ebase = w.g_off[s] * CTX_S;
        }
        if (j < total) {
          const real tq = *tq_ptr;
          const real th_ = (tq - r_tp) * r_inv;
#pragma unroll
          for (int i = 0; i < CTX_N; ++i) {
            real v = r_coef[CTX_DEG * CTX_N + i];
#pragma unroll
            for (int mm = CTX_DEG - 1; mm >= 0; --mm) v = v * th_ + r_coef[mm * CTX_N + i];
            if (i < CTX_S) w.stage_y[lane * CTX_S + i] = v;
#if CTX_TAN
            else {
              const int dd = (i - CTX_S) / CTX_S, ii = (i - CTX_S) - dd * CTX_S;
              w.stage_s[lane * (CTX_S * CTX_D) + ii * CTX_D + dd] = v;
            }
#endif
          }
          if (!uniform) w.gpt[lane] = r_ebase + (long long)j * CTX_S;
        }
        tq_ptr += 32;
        __syncwarp();
        if (uniform) {
          // one contiguous run of a single trajectory: elements ebase + r0*S ...
          real* dst = a.ys + (r_ebase + (long long)r0 * CTX_S);
#pragma unroll
          for (int mm = 0; mm < CTX_S; ++mm) {
            const int e = lane + 32 * mm;
            if (e < cnt * CTX_S) dst[e] = w.stage_y[e];
          }
#if CTX_TAN
          if (a.sens) {
            real* ds = a.sens + (r_ebase + (long long)r0 * CTX_S) * CTX_D;
#pragma unroll
            for (int mm = 0; mm < CTX_S * CTX_D; ++mm) {
              const int e = lane + 32 * mm;
              if (e < cnt * (CTX_S * CTX_D)) ds[e] = w.stage_s[e];
            }
          }
#endif
        } else {
#pragma unroll
          for (int mm = 0; mm < CTX_S; ++mm) {
            const int e = lane + 32 * mm;
            if (e < cnt * CTX_S) {
              const int p = e / CTX_S, i = e - p * CTX_S;
              a.ys[w.gpt[p] + i] = w.stage_y[e];
            }
          }
#if CTX_TAN
          if (a.sens) {
#pragma unroll
            for (int mm = 0; mm < CTX_S * CTX_D; ++mm) {
              const int e = lane + 32 * mm;
              if (e < cnt * (CTX_S * CTX_D)) {
                const int p = e / (CTX_S * CTX_D), r = e - p * (CTX_S * CTX_D);
                a.sens[w.gpt[p] * CTX_D + r] = w.stage_s[e];
              }
            }
          }
#endif
        }
        __syncwarp();
      }
      }  // total > 0
    }

#endif  // CTX_TAN

    // ------------------------------------------------------------- advance / retire
    if (accepted) {
      idx += nsave;
#pragma unroll
      for (int i = 0; i < CTX_N; ++i) y[i] = ynew[i];
#if CTX_SOLVER == 0 || CTX_SOLVER == 1
#pragma unroll
      for (int i = 0; i < CTX_N; ++i) K[0][i] = K[6][i];
#else
      ctx_field(tprev, y, th, c, y0, K[0]);
#endif
    }
    if ((retire && phase == 0) || filling) {
      a.status[b] = status;
      a.n_accepted[b] = nacc;
      a.n_rejected[b] = nrej;
      b = -1;
      phase = 0;
    }
  }
#if CTX_BULK && !CTX_TAN
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores have landed
#endif
}

extern "C" __global__ void __launch_bounds__(CTX_WARPS * 32, CTX_MIN_BLOCKS)
ctx_solve_v1(const ctx_solve_args a) {
  if (a.ts_smem) ctx_solve_v1_body<true>(a);
  else ctx_solve_v1_body<false>(a);
}

// bytes of dynamic shared memory one warp of ctx_solve_v1 needs (read by the host at load time)
extern "C" __device__ const unsigned ctx_v1_warp_smem_bytes = sizeof(ctx_warp_smem);

#if !CTX_TAN && !CTX_THETA_PTR && CTX_QUEUE && (CTX_SOLVER == 0 || CTX_SOLVER == 1) && !CTX_LLSAVE
#define CTX_HAVE_V2 1
#include "ctx_solve_v2.cuh"
#endif

#endif  // CTX_NO_V1

// =====================================================================================
// kernel K2: fused solve + forward tangents + masked log-likelihood and its gradient
// =====================================================================================
// Replaces, per NUTS leapfrog step, value_and_grad of the numpyro model
//   catalax/mcmc/mcmc.py:405-491   states = sim_func(y0s, theta, constants, times)   (:448)
//                                  loc = states[..., observables]                    (:461)
//                                  sample("y", likelihood(loc, sqrt(sigma_y**2)), obs=data)
//                                  under handlers.mask(mask)                         (:486-491)
// One trajectory per (chain, series) pair, u = chain*M + m: theta and sigma are per chain,
// y0 / constants / times / data / mask per series (in_axes=(0, None, 0, 0), mcmc.py:860).
// Nothing of size [M,T,S] is written: every lane evaluates its own trajectory's requested
// times right after the accepted step that covers them and accumulates
//   ll, d ll/d theta_p = sum dlogp/dloc * d y_o/d theta_p, d ll/d sigma_o, d ll/d y0_i
// in registers; one row of partial[u, :] is written when the trajectory retires and
// ctx_loglik_reduce sums the M rows of every chain with warp shuffles (deterministic).
// observables = all modeled states in state order (mcmc.py:393-395).
//   likelihood 0: Normal       log p = -z^2/2 - log s - log(2 pi)/2
//   likelihood 1: SoftLaplace  log p = log(2/pi) - log s - logaddexp(z, -z)   (numpyro)
//   z = (x - loc)/s,  s = sqrt(sigma^2) = |sigma|
#if CTX_TAN && !CTX_THETA_PTR
#define CTX_KP (1 + CTX_D_THETA + CTX_S + CTX_D_Y0)

struct ctx_loglik_args {
  const real* y0s;      // [M,S]
  const real* theta;    // [CH,P]
  const real* consts;   // [M,C]
  const real* ts;       // [M,T]
  const real* data;     // [M,T,S]
  const unsigned char* mask;  // [M,T,S] 1 = observed; null -> isfinite(data)
  const real* sigma;    // [CH,S]
  real* partial;        // [CH*M, KP] = (ll, d/dtheta[D_THETA], d/dsigma[S], d/dy0[D_Y0])
  int* status;          // [CH*M]
  int* n_accepted;      // [CH*M]
  int* n_rejected;      // [CH*M]
  unsigned long long* work_counter;
  long long n_traj;     // CH*M
  int M, T, likelihood, max_steps;
  real rtol, atol, dt0;
  const int* order;     // scheduling hint: position u of the work queue is trajectory order[u]; null = u
};

__device__ __forceinline__ real ctx_log1p_exp_neg2a(const real a) {  // log(1 + exp(-2a)), a >= 0
  return log1p(exp(R(-2.0) * a));
}

// Register cap of the log-likelihood kernels.  Uncapped they take 164 (f32) registers = 12 warps
// per SM; capped to 4 CTAs/SM (128 registers, a few spilled words) the f32 kernel of config 4 runs
// 1.92 -> 1.69 ms, f64 at 3 CTAs/SM 6.43 -> 6.20 ms, same bits (measured on B200, round 2); 5 CTAs/SM
// loses (2.12 ms).  The host passes the cap per model (ctx_api.cu compile_variant).
#ifndef CTX_LL_MIN_BLOCKS
#define CTX_LL_MIN_BLOCKS CTX_MIN_BLOCKS
#endif

extern "C" __global__ void __launch_bounds__(CTX_WARPS * 32, CTX_LL_MIN_BLOCKS)
ctx_loglik_v1(const ctx_loglik_args a) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int T = a.T;

  long long u = -1;
  real y[CTX_N], ynew[CTX_N], y0[CTX_S];
  real th[CTX_P > 0 ? CTX_P : 1], c[CTX_C > 0 ? CTX_C : 1], sg[CTX_S];
  real K[CTX_NST][CTX_N];
  real acc[CTX_KP];
  real tprev = R(0.0), tnext = R(0.0), t_end = R(0.0);
  int idx = 0, nacc = 0, nrej = 0, status = 0, m = 0;
  bool exhausted = false;
  const real* tsr = a.ts;

  while (true) {
    const unsigned need = __ballot_sync(CTX_FULL, u < 0);
    if (need && !exhausted) {
      const int cnt = __popc(need);
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.work_counter, (unsigned long long)cnt);
      base = __shfl_sync(CTX_FULL, base, 0);
      if (base + cnt >= (unsigned long long)a.n_traj) exhausted = true;
      if (u < 0) {
        const unsigned long long nu = base + __popc(need & lt_mask);
        if (nu < (unsigned long long)a.n_traj) {
          u = a.order ? (long long)a.order[nu] : (long long)nu;
          const long long ch = u / a.M;
          m = (int)(u - ch * a.M);
#pragma unroll
          for (int i = 0; i < CTX_S; ++i) y0[i] = y[i] = a.y0s[(size_t)m * CTX_S + i];
#pragma unroll
          for (int i = 0; i < CTX_P; ++i) th[i] = a.theta[(size_t)ch * CTX_P + i];
#pragma unroll
          for (int i = 0; i < CTX_C; ++i) c[i] = a.consts[(size_t)m * CTX_C + i];
#pragma unroll
          for (int i = 0; i < CTX_S; ++i) sg[i] = a.sigma[(size_t)ch * CTX_S + i];
#pragma unroll
          for (int i = CTX_S; i < CTX_N; ++i) y[i] = R(0.0);
#pragma unroll
          for (int i = 0; i < CTX_D_Y0; ++i) y[CTX_S + (CTX_D_THETA + i) * CTX_S + i] = R(1.0);
#pragma unroll
          for (int k = 0; k < CTX_KP; ++k) acc[k] = R(0.0);
          tsr = a.ts + (size_t)m * T;
          t_end = tsr[T - 1];
          tprev = R(0.0);
          tnext = fmin(tprev + a.dt0, t_end);
          idx = 0; nacc = 0; nrej = 0; status = 0;
          ctx_field(tprev, y, th, c, y0, K[0]);
        }
      }
    }
    if (__ballot_sync(CTX_FULL, u >= 0) == 0u) break;

    if (u >= 0) {
      bool keep = true, done = !(tprev < t_end);
      const real tp = tprev, tn_step = tnext;
      if (!done) {
        const real dt = tnext - tprev;
        ctx_rk_stages(tprev, dt, y, K, th, c, y0, ynew);
        real dt_next = a.dt0;
#if CTX_ADAPTIVE
        bool bad_norm = false;
        real factor = ctx_error_control(dt, y, ynew, K, a.rtol, a.atol, keep, bad_norm);
        factor = fmin(fmax(factor, keep ? R(1.0) : R(0.2)), R(10.0));
        dt_next = dt * factor;
        if (bad_norm) dt_next = CTX_QNAN;
#endif
        real tp_new = keep ? tnext : tprev;
        if (keep) ++nacc; else ++nrej;
        tp_new = fmin(tp_new, t_end);
        real tn = tp_new + dt_next;
        if (tn > t_end - CTX_TOL_END) tn = keep ? t_end : tp_new + R(0.5) * (t_end - tp_new);
        done = !(tp_new < t_end);
        if (!done) {
          if (!(fabs(tn) < CTX_INF)) status = CTX_ST_BAD;
          else if (nacc + nrej >= a.max_steps) status = CTX_ST_MAX;
        }
        tprev = tp_new; tnext = tn;
      } else {
        // t0 == t1: every requested time is t0; interpolate a zero-length step at theta = 0
#pragma unroll
        for (int i = 0; i < CTX_N; ++i) ynew[i] = y[i];
#pragma unroll
        for (int k = 1; k < CTX_NST; ++k)
#pragma unroll
          for (int i = 0; i < CTX_N; ++i) K[k][i] = R(0.0);
      }
      // ---- requested times covered by this accepted step: likelihood terms
      if (keep) {
        while (idx < T) {
          const real tq = tsr[idx];
          if (!done && !(tq <= tn_step)) break;
          real out[CTX_N];
          ctx_dense(tq, tp, done && tp == tn_step ? tp : tn_step, y, ynew, K, out);
          const real* drow = a.data + ((size_t)m * T + idx) * CTX_S;
          const unsigned char* mrow = a.mask ? a.mask + ((size_t)m * T + idx) * CTX_S : nullptr;
#pragma unroll
          for (int o = 0; o < CTX_S; ++o) {
            const real x = drow[o];
            const bool use = mrow ? (mrow[o] != 0) : (fabs(x) < CTX_INF);
            if (use) {
              const real s = fabs(sg[o]);
              const real z = (x - out[o]) / s;
              real lp, dl_dloc, dl_ds;
              if (a.likelihood == 0) {
                lp = R(-0.5) * z * z - log(s) - R(0.9189385332046727418);
                dl_dloc = z / s;
                dl_ds = (z * z - R(1.0)) / s;
              } else {
                const real az = fabs(z);
                lp = R(-0.4515827052894548647) - log(s) - (az + ctx_log1p_exp_neg2a(az));
                const real tz = tanh(z);
                dl_dloc = tz / s;
                dl_ds = (z * tz - R(1.0)) / s;
              }
              acc[0] += lp;
#pragma unroll
              for (int d = 0; d < CTX_D_THETA; ++d) acc[1 + d] += dl_dloc * out[CTX_S + d * CTX_S + o];
              acc[1 + CTX_D_THETA + o] += dl_ds * (sg[o] < R(0.0) ? R(-1.0) : R(1.0));
#pragma unroll
              for (int d = 0; d < CTX_D_Y0; ++d)
                acc[1 + CTX_D_THETA + CTX_S + d] += dl_dloc * out[CTX_S + (CTX_D_THETA + d) * CTX_S + o];
            }
          }
          ++idx;
        }
#pragma unroll
        for (int i = 0; i < CTX_N; ++i) y[i] = ynew[i];
#if CTX_SOLVER == 0 || CTX_SOLVER == 1
#pragma unroll
        for (int i = 0; i < CTX_N; ++i) K[0][i] = K[6][i];
#else
        ctx_field(tprev, y, th, c, y0, K[0]);
#endif
      }
      if (done || status != 0) {
        if (status != 0) {
          // reference: unsaved states are inf (constant) -> log p = -inf, gradient inf * 0 = NaN
          bool any_left = false;
          for (int j = idx; j < T && !any_left; ++j)
            for (int o = 0; o < CTX_S; ++o) {
              const real x = a.data[((size_t)m * T + j) * CTX_S + o];
              const bool use = a.mask ? (a.mask[((size_t)m * T + j) * CTX_S + o] != 0) : (fabs(x) < CTX_INF);
              any_left |= use;
            }
          if (any_left) {
            acc[0] = -CTX_INF;
#pragma unroll
            for (int k = 1; k < CTX_KP; ++k) acc[k] = CTX_INF - CTX_INF;
          }
        }
#pragma unroll
        for (int k = 0; k < CTX_KP; ++k) a.partial[(size_t)u * CTX_KP + k] = acc[k];
        a.status[u] = status;
        a.n_accepted[u] = nacc;
        a.n_rejected[u] = nrej;
        u = -1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// ctx_loglik_v2: the same computation with WARP-COOPERATIVE likelihood terms.
// In ctx_loglik_v1 every lane evaluates the requested times of its own accepted step in a
// data-dependent loop; a step covers ~1 of the 20 requested times on average, so that loop runs
// with a handful of active lanes (ncu: 10 of 32 threads per instruction over the whole kernel).
// Here a lane with an accepted step publishes the step's interpolant (polynomial coefficients of
// all S*(1+D) components) in shared memory; the requested times covered by all lanes' steps are
// flattened and dealt out one per lane (full warps), each lane evaluates one time point: Horner,
// the likelihood of the S observables and their gradient contributions.  The contributions are
// added to the owning trajectory's accumulators (shared memory) by the last lane of every
// segment, sequentially in time order starting from the accumulator's value: the summation order
// of a trajectory is independent of which lanes evaluated it, so results are reproducible.
// tp, 1/dt, coefficients [power][component], then per observable 1/|sigma|, 1/sigma, log|sigma|
// (the chain's scale enters every term only through these: computed once by the publisher)
#define CTX_LL_RW (2 + CTX_NCOEF + 3 * CTX_S)
// Optional input queue of the cooperative kernel (as in ctx_solve_v1: every warp claims 32
// trajectories at a time and copies their operands with cp.async into one of two shared-memory
// buffers).  Built and measured on B200 (config 4, round 2) -- and it LOSES: under the 128-register
// cap its 64-bit bookkeeping spills (f32 24 B, f64 272 B of stack): f32 1.68 -> 1.89 ms, with the
// order hint 1.28 -> 1.44 ms, f64 6.21 -> 8.55 ms.  The direct refill (atomic claim + operand loads
// by the retiring lanes) stays the default; its stalls are what the order hint removes (lanes of a
// warp retire together).
#ifndef CTX_LL_QUEUE
#define CTX_LL_QUEUE 0
#endif
#define CTX_LLQW (2 * CTX_S + CTX_P + CTX_C + 1)   // y0, theta, constants, sigma, ts[T-1]
struct ctx_ll_warp_smem {
  real rec[CTX_LL_RW][32];   // one column per publishing lane of this iteration
  real acc[CTX_KP][32];      // accumulators of the trajectory each lane integrates
  int pt0[32];               // point index (m*T + idx) of the segment's first time - its flat offset
  int off[32];               // flat offset of the segment
  int cnt[32];               // requested times it covers
  int owner[32];             // publishing lane
#if CTX_LL_QUEUE
  real q_in[2][CTX_LLQW][32];   // input queue: operands of the next trajectories (two buffers)
  int q_idx[2][32];             // their trajectory indices (order hint applied)
#endif
};

#if CTX_LL_QUEUE
// operands of the 32 queue positions [base, base + 32) -> buffer `buf` (cp.async, lane = entry)
__device__ __forceinline__ void ctx_ll_queue_fetch(const ctx_loglik_args& a, ctx_ll_warp_smem& w,
                                                   const int buf, const unsigned long long base,
                                                   const int lane) {
  const unsigned long long uq = base + (unsigned long long)lane;
  if (uq < (unsigned long long)a.n_traj) {
    const long long u = a.order ? (long long)a.order[uq] : (long long)uq;
    w.q_idx[buf][lane] = (int)u;
    const long long ch = u / a.M;
    const long long m = u - ch * a.M;
    int k = 0;
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) ctx_cp_async(&w.q_in[buf][k++][lane], a.y0s + (size_t)m * CTX_S + i);
#pragma unroll
    for (int i = 0; i < CTX_P; ++i) ctx_cp_async(&w.q_in[buf][k++][lane], a.theta + (size_t)ch * CTX_P + i);
#pragma unroll
    for (int i = 0; i < CTX_C; ++i) ctx_cp_async(&w.q_in[buf][k++][lane], a.consts + (size_t)m * CTX_C + i);
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) ctx_cp_async(&w.q_in[buf][k++][lane], a.sigma + (size_t)ch * CTX_S + i);
    ctx_cp_async(&w.q_in[buf][k++][lane], a.ts + (size_t)m * a.T + (a.T - 1));
  }
  CTX_CP_COMMIT();
}
#endif

extern "C" __global__ void __launch_bounds__(CTX_WARPS * 32, CTX_LL_MIN_BLOCKS)
ctx_loglik_v2(const ctx_loglik_args a) {
  extern __shared__ __align__(16) unsigned char ctx_dyn_smem[];
  ctx_ll_warp_smem& w = reinterpret_cast<ctx_ll_warp_smem*>(ctx_dyn_smem)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned le_mask = CTX_FULL >> (31 - lane);
  const int T = a.T;

  long long u = -1;
  real y[CTX_N], ynew[CTX_N], y0[CTX_S];
  real th[CTX_P > 0 ? CTX_P : 1], c[CTX_C > 0 ? CTX_C : 1], sg[CTX_S];
  real K[CTX_NST][CTX_N];
  real tprev = R(0.0), tnext = R(0.0), t_end = R(0.0);
  int idx = 0, nacc = 0, nrej = 0, status = 0, m = 0;
  const real* tsr = a.ts;
#if CTX_LL_QUEUE
  const unsigned long long NT = (unsigned long long)a.n_traj;
  unsigned long long qb_cur = 0, qb_nxt = 0, qb_pend = 0;
  int q_cur = 0, q_pos = 0;
  bool q_empty = false;
  {
    if (lane == 0) qb_cur = atomicAdd(a.work_counter, 64ull);
    qb_cur = __shfl_sync(CTX_FULL, qb_cur, 0);
    qb_nxt = qb_cur + 32ull;
    ctx_ll_queue_fetch(a, w, 0, qb_cur, lane);
    ctx_ll_queue_fetch(a, w, 1, qb_nxt, lane);
    if (lane == 0) qb_pend = atomicAdd(a.work_counter, 32ull);   // read one buffer later
    CTX_CP_WAIT(1);
    __syncwarp();
    q_empty = qb_cur >= NT;
  }

  while (true) {
    const unsigned need = __ballot_sync(CTX_FULL, u < 0);
    if (need && !q_empty) {
      const int nval = (int)min(32ull, NT - qb_cur);
      const int take = min(__popc(need), nval - q_pos);
      const int rank = __popc(need & lt_mask);
      if (u < 0 && rank < take) {
        const int e = q_pos + rank;
        u = (long long)w.q_idx[q_cur][e];
        const long long ch = u / a.M;
        m = (int)(u - ch * a.M);
        int k = 0;
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) y0[i] = y[i] = w.q_in[q_cur][k++][e];
#pragma unroll
        for (int i = 0; i < CTX_P; ++i) th[i] = w.q_in[q_cur][k++][e];
#pragma unroll
        for (int i = 0; i < CTX_C; ++i) c[i] = w.q_in[q_cur][k++][e];
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) sg[i] = w.q_in[q_cur][k++][e];
        t_end = w.q_in[q_cur][k++][e];
#pragma unroll
        for (int i = CTX_S; i < CTX_N; ++i) y[i] = R(0.0);
#pragma unroll
        for (int i = 0; i < CTX_D_Y0; ++i) y[CTX_S + (CTX_D_THETA + i) * CTX_S + i] = R(1.0);
#pragma unroll
        for (int kk = 0; kk < CTX_KP; ++kk) w.acc[kk][lane] = R(0.0);
        tsr = a.ts + (size_t)m * T;
        tprev = R(0.0);
        tnext = fmin(tprev + a.dt0, t_end);
        idx = 0; nacc = 0; nrej = 0; status = 0;
        ctx_field(tprev, y, th, c, y0, K[0]);
      }
      q_pos += take;
      if (q_pos >= nval) {
        CTX_CP_WAIT(0);
        __syncwarp();
        const unsigned long long nbase = __shfl_sync(CTX_FULL, qb_pend, 0);
        ctx_ll_queue_fetch(a, w, q_cur, nbase, lane);
        qb_cur = qb_nxt;
        qb_nxt = nbase;
        if (lane == 0 && nbase < NT) qb_pend = atomicAdd(a.work_counter, 32ull);
        q_cur ^= 1; q_pos = 0;
        q_empty = qb_cur >= NT;
      }
    }
    if (__ballot_sync(CTX_FULL, u >= 0) == 0u) {
      if (q_empty) break;
      continue;
    }
#else
  bool exhausted = false;

  while (true) {
    const unsigned need = __ballot_sync(CTX_FULL, u < 0);
    if (need && !exhausted) {
      const int cnt = __popc(need);
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.work_counter, (unsigned long long)cnt);
      base = __shfl_sync(CTX_FULL, base, 0);
      if (base + cnt >= (unsigned long long)a.n_traj) exhausted = true;
      if (u < 0) {
        const unsigned long long nu = base + __popc(need & lt_mask);
        if (nu < (unsigned long long)a.n_traj) {
          u = a.order ? (long long)a.order[nu] : (long long)nu;
          const long long ch = u / a.M;
          m = (int)(u - ch * a.M);
#pragma unroll
          for (int i = 0; i < CTX_S; ++i) y0[i] = y[i] = a.y0s[(size_t)m * CTX_S + i];
#pragma unroll
          for (int i = 0; i < CTX_P; ++i) th[i] = a.theta[(size_t)ch * CTX_P + i];
#pragma unroll
          for (int i = 0; i < CTX_C; ++i) c[i] = a.consts[(size_t)m * CTX_C + i];
#pragma unroll
          for (int i = 0; i < CTX_S; ++i) sg[i] = a.sigma[(size_t)ch * CTX_S + i];
#pragma unroll
          for (int i = CTX_S; i < CTX_N; ++i) y[i] = R(0.0);
#pragma unroll
          for (int i = 0; i < CTX_D_Y0; ++i) y[CTX_S + (CTX_D_THETA + i) * CTX_S + i] = R(1.0);
#pragma unroll
          for (int k = 0; k < CTX_KP; ++k) w.acc[k][lane] = R(0.0);
          tsr = a.ts + (size_t)m * T;
          t_end = tsr[T - 1];
          tprev = R(0.0);
          tnext = fmin(tprev + a.dt0, t_end);
          idx = 0; nacc = 0; nrej = 0; status = 0;
          ctx_field(tprev, y, th, c, y0, K[0]);
        }
      }
    }
    if (__ballot_sync(CTX_FULL, u >= 0) == 0u) break;
#endif   // CTX_LL_QUEUE

    // ------------------------------------------------------------- one step attempt
    bool keep = false, done = false;
    int nsave = 0;
    real rec_tp = R(0.0), rec_dt = R(0.0);
    if (u >= 0) {
      keep = true;
      done = !(tprev < t_end);
      rec_tp = tprev;
      if (!done) {
        const real dt = tnext - tprev;
        const real tn_step = tnext;
        rec_dt = dt;
        ctx_rk_stages(tprev, dt, y, K, th, c, y0, ynew);
        real dt_next = a.dt0;
#if CTX_ADAPTIVE
        bool bad_norm = false;
        real factor = ctx_error_control(dt, y, ynew, K, a.rtol, a.atol, keep, bad_norm);
        factor = fmin(fmax(factor, keep ? R(1.0) : R(0.2)), R(10.0));
        dt_next = dt * factor;
        if (bad_norm) dt_next = CTX_QNAN;
#endif
        real tp_new = keep ? tnext : tprev;
        if (keep) ++nacc; else ++nrej;
        tp_new = fmin(tp_new, t_end);
        real tn = tp_new + dt_next;
        if (tn > t_end - CTX_TOL_END) tn = keep ? t_end : tp_new + R(0.5) * (t_end - tp_new);
        done = !(tp_new < t_end);
        if (!done) {
          if (!(fabs(tn) < CTX_INF)) status = CTX_ST_BAD;
          else if (nacc + nrej >= a.max_steps) status = CTX_ST_MAX;
        }
        tprev = tp_new; tnext = tn;
        if (keep)
          nsave = done ? T - idx : ctx_upper_bound<false>(tsr, 0u, idx, T, tn_step, t_end) - idx;
      } else {
        // t0 == t1: every requested time is t0 (a zero-length step evaluated at theta = 0)
#pragma unroll
        for (int i = 0; i < CTX_N; ++i) ynew[i] = y[i];
#pragma unroll
        for (int k = 1; k < CTX_NST; ++k)
#pragma unroll
          for (int i = 0; i < CTX_N; ++i) K[k][i] = R(0.0);
        nsave = T - idx;
      }
    }

    // ------------------------------------------------------------- cooperative likelihood terms
    const unsigned nz = __ballot_sync(CTX_FULL, nsave > 0);
    if (nz) {
      int incl = nsave;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(CTX_FULL, incl, d);
        if (lane >= d) incl += v;
      }
      const int off = incl - nsave;
      const int total = __shfl_sync(CTX_FULL, incl, 31);
      if (nsave > 0) {
        const int s = __popc(nz & lt_mask);
        real coef[CTX_NCOEF];
        ctx_dense_coefs(rec_dt, y, ynew, K, coef);
        w.rec[0][s] = rec_tp;
        w.rec[1][s] = (rec_dt == R(0.0)) ? R(0.0) : CTX_IDIV(R(1.0), rec_dt);
#pragma unroll
        for (int i = 0; i < CTX_NCOEF; ++i) w.rec[2 + i][s] = coef[i];
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) {
          const real sa_ = fabs(sg[i]);
          w.rec[2 + CTX_NCOEF + 3 * i][s] = R(1.0) / sa_;
          w.rec[2 + CTX_NCOEF + 3 * i + 1][s] = R(1.0) / sg[i];
          w.rec[2 + CTX_NCOEF + 3 * i + 2][s] = log(sa_);
        }
        w.pt0[s] = m * T + idx - off;
        w.off[s] = off;
        w.cnt[s] = nsave;
        w.owner[s] = lane;
      }
      __syncwarp();
      const unsigned myoff = nsave > 0 ? (unsigned)off : 0xffffffffu;
      int seg_base = 0;
      for (int r0 = 0; r0 < total; r0 += 32) {
        const unsigned dpos = myoff - (unsigned)r0;
        const unsigned bits = __reduce_or_sync(CTX_FULL, dpos < 32u ? (1u << dpos) : 0u);
        const int s = seg_base + __popc(bits & le_mask) - 1;   // >= 0: flat position 0 starts one
        seg_base += __popc(bits);
        const int j = r0 + lane;
        const bool act = j < total;
        real cv[CTX_KP];
#pragma unroll
        for (int k = 0; k < CTX_KP; ++k) cv[k] = R(0.0);
        int my_len = 0, s_owner = 0;
        bool is_tail = false;
        if (act) {
          const int s_off = w.off[s], s_cnt = w.cnt[s];
          s_owner = w.owner[s];
          const int first = max(s_off - r0, 0);                 // this segment's lanes in the pass
          const int last = min(s_off + s_cnt - r0, 32) - 1;
          is_tail = lane == last;
          my_len = last - first + 1;
          const int pt = w.pt0[s] + j;
          const real tq = a.ts[pt];
          const real th_ = (tq - w.rec[0][s]) * w.rec[1][s];
          real out[CTX_N];
#pragma unroll
          for (int i = 0; i < CTX_N; ++i) {
            real v = w.rec[2 + CTX_DEG * CTX_N + i][s];
#pragma unroll
            for (int mm = CTX_DEG - 1; mm >= 0; --mm) v = v * th_ + w.rec[2 + mm * CTX_N + i][s];
            out[i] = v;
          }
          const real* drow = a.data + (size_t)pt * CTX_S;
          const unsigned char* mrow = a.mask ? a.mask + (size_t)pt * CTX_S : nullptr;
#pragma unroll
          for (int o = 0; o < CTX_S; ++o) {
            const real x = drow[o];
            const bool use = mrow ? (mrow[o] != 0) : (fabs(x) < CTX_INF);
            if (use) {
              const real inv_sd = w.rec[2 + CTX_NCOEF + 3 * o][s];       // 1 / |sigma|
              const real inv_sg = w.rec[2 + CTX_NCOEF + 3 * o + 1][s];   // 1 / sigma (sign of d|s|/ds)
              const real log_sd = w.rec[2 + CTX_NCOEF + 3 * o + 2][s];
              const real z = (x - out[o]) * inv_sd;
              real lp, dl_dloc, dl_ds;   // dl_ds: d lp / d sigma (sign included)
              if (a.likelihood == 0) {
                lp = R(-0.5) * z * z - log_sd - R(0.9189385332046727418);
                dl_dloc = z * inv_sd;
                dl_ds = (z * z - R(1.0)) * inv_sg;
              } else {
                // one expm1 serves both log(1 + exp(-2|z|)) and tanh(|z|) = -em1 / (2 + em1)
                const real az = fabs(z);
                const real em1 = expm1(R(-2.0) * az);
                lp = R(-0.4515827052894548647) - log_sd - (az + log1p(R(1.0) + em1));
                const real ta = -em1 / (R(2.0) + em1);
                const real tz = z < R(0.0) ? -ta : ta;
                dl_dloc = tz * inv_sd;
                dl_ds = (z * tz - R(1.0)) * inv_sg;
              }
              cv[0] += lp;
#pragma unroll
              for (int d = 0; d < CTX_D_THETA; ++d) cv[1 + d] += dl_dloc * out[CTX_S + d * CTX_S + o];
              cv[1 + CTX_D_THETA + o] += dl_ds;
#pragma unroll
              for (int d = 0; d < CTX_D_Y0; ++d)
                cv[1 + CTX_D_THETA + CTX_S + d] += dl_dloc * out[CTX_S + (CTX_D_THETA + d) * CTX_S + o];
            }
          }
        }
        // the last lane of a segment adds the segment's terms to the owner's accumulators in
        // time order: acc = (((acc + c_first) + ...) + c_last)
        const int maxlen = __reduce_max_sync(CTX_FULL, my_len);
        real av[CTX_KP];
        if (is_tail) {
#pragma unroll
          for (int k = 0; k < CTX_KP; ++k) av[k] = w.acc[k][s_owner];
        }
        for (int q = maxlen - 1; q >= 0; --q) {
#pragma unroll
          for (int k = 0; k < CTX_KP; ++k) {
            const real t = __shfl_sync(CTX_FULL, cv[k], (lane - q) & 31);
            if (is_tail && q < my_len) av[k] += t;
          }
        }
        if (is_tail) {
#pragma unroll
          for (int k = 0; k < CTX_KP; ++k) w.acc[k][s_owner] = av[k];
        }
        __syncwarp();   // a segment that continues in the next pass reads what this one wrote
      }
    }

    // ------------------------------------------------------------- advance / retire
    if (u >= 0) {
      if (keep) {
        idx += nsave;
#pragma unroll
        for (int i = 0; i < CTX_N; ++i) y[i] = ynew[i];
#if CTX_SOLVER == 0 || CTX_SOLVER == 1
#pragma unroll
        for (int i = 0; i < CTX_N; ++i) K[0][i] = K[6][i];
#else
        ctx_field(tprev, y, th, c, y0, K[0]);
#endif
      }
      if (done || status != 0) {
        real acc[CTX_KP];
#pragma unroll
        for (int k = 0; k < CTX_KP; ++k) acc[k] = w.acc[k][lane];
        if (status != 0) {
          // reference: unsaved states are inf (constant) -> log p = -inf, gradient inf * 0 = NaN
          bool any_left = false;
          for (int j = idx; j < T && !any_left; ++j)
            for (int o = 0; o < CTX_S; ++o) {
              const real x = a.data[((size_t)m * T + j) * CTX_S + o];
              const bool use = a.mask ? (a.mask[((size_t)m * T + j) * CTX_S + o] != 0) : (fabs(x) < CTX_INF);
              any_left |= use;
            }
          if (any_left) {
            acc[0] = -CTX_INF;
#pragma unroll
            for (int k = 1; k < CTX_KP; ++k) acc[k] = CTX_INF - CTX_INF;
          }
        }
#pragma unroll
        for (int k = 0; k < CTX_KP; ++k) a.partial[(size_t)u * CTX_KP + k] = acc[k];
        a.status[u] = status;
        a.n_accepted[u] = nacc;
        a.n_rejected[u] = nrej;
        u = -1;
      }
    }
  }
}

extern "C" __device__ const unsigned ctx_ll_warp_smem_bytes = sizeof(ctx_ll_warp_smem);

// out[ch, k] = sum_m partial[ch*M + m, k] for k < KR; one warp per chain, double accumulation.
extern "C" __global__ void __launch_bounds__(128)
ctx_loglik_reduce(const real* partial, real* out, const long long n_chains, const int M,
                  const int KR) {
  const long long ch = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ch >= n_chains) return;
  for (int k = 0; k < KR; ++k) {
    double s = 0.0;
    for (int mm = lane; mm < M; mm += 32) s += (double)partial[((size_t)ch * M + mm) * CTX_KP + k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(CTX_FULL, s, d);
    if (lane == 0) out[(size_t)ch * KR + k] = (real)s;
  }
}

// =====================================================================================
// kernel K5: surrogate (rate-matching) log-likelihood + gradient -- no solve
// =====================================================================================
// Replaces value_and_grad of the numpyro model in Modes.SURROGATE:
//   catalax/mcmc/mcmc.py:819-847  sim_func = vmap(rate_fun)(y_points, theta, constants, t, inits)
//                                 (measured states as evaluation points, "{state}_init" from the
//                                 first time of each measurement, data = surrogate rates)
//   catalax/mcmc/mcmc.py:469-480  sigma_total[p,j] = sum_k J[p,j,k]^2 sigma_y[k]^2
//                                 (+ sigma_surr^2 + rate_sigma^2) + 1e-8
//   catalax/mcmc/mcmc.py:486-491  sample("y", likelihood(loc, sqrt(sigma_total)), obs=data) | mask
// One warp per chain; lanes stride over the N = M * n_time points, evaluate the generated field
// and its parameter derivatives (ctx_rhs_tan with zero tangents: ds = df/dtheta), accumulate
//   (ll, d ll/d theta[P], d ll/d sigma_y[S])
// and reduce with warp shuffles in double precision (fixed order: reproducible).
struct ctx_rate_ll_args {
  const real* t;        // [N]
  const real* y;        // [N,S]     evaluation points (measured states), p = m * n_time + j
  const real* theta;    // [CH,P]
  const real* consts;   // [M,C]     row p / n_time
  const real* inits;    // [M,S]     "{state}_init" values, row p / n_time
  const real* data;     // [N,S]     surrogate rates
  const unsigned char* mask;  // [N,S] or null -> isfinite(data)
  const real* jac2;     // [N,S,S]   squared surrogate rate-Jacobian J[p,j,k]^2
  const real* extra2;   // [N,S] or null: sigma_surr^2 + rate_sigma^2
  const real* sigma;    // [CH,S]
  real* out;            // [CH, 1+P+S]
  long long n_chains;
  int N, n_time, likelihood;
};
#define CTX_KR (1 + CTX_D_THETA + CTX_S)

extern "C" __global__ void __launch_bounds__(128) ctx_rate_loglik(const ctx_rate_ll_args a) {
  const long long ch = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ch >= a.n_chains) return;
  real th[CTX_P > 0 ? CTX_P : 1], sg[CTX_S], acc[CTX_KR];
#pragma unroll
  for (int i = 0; i < CTX_P; ++i) th[i] = a.theta[(size_t)ch * CTX_P + i];
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) sg[i] = a.sigma[(size_t)ch * CTX_S + i];
#pragma unroll
  for (int k = 0; k < CTX_KR; ++k) acc[k] = R(0.0);
  for (int p = lane; p < a.N; p += 32) {
    const int mrow = p / a.n_time;
    real y[CTX_S], y0[CTX_S], c[CTX_C > 0 ? CTX_C : 1], dy[CTX_S];
    real s[CTX_S * CTX_D], ds[CTX_S * CTX_D];
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) y[i] = a.y[(size_t)p * CTX_S + i];
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) y0[i] = a.inits[(size_t)mrow * CTX_S + i];
#pragma unroll
    for (int i = 0; i < CTX_C; ++i) c[i] = a.consts[(size_t)mrow * CTX_C + i];
#pragma unroll
    for (int i = 0; i < CTX_S * CTX_D; ++i) s[i] = R(0.0);
    ctx_rhs_tan(a.t[p], y, s, th, c, y0, dy, ds);
#pragma unroll
    for (int j = 0; j < CTX_S; ++j) {
      const real x = a.data[(size_t)p * CTX_S + j];
      const bool use = a.mask ? (a.mask[(size_t)p * CTX_S + j] != 0) : (fabs(x) < CTX_INF);
      if (use) {
        real j2[CTX_S];
        real s2 = a.extra2 ? a.extra2[(size_t)p * CTX_S + j] : R(0.0);
#pragma unroll
        for (int k = 0; k < CTX_S; ++k) {
          j2[k] = a.jac2[((size_t)p * CTX_S + j) * CTX_S + k];
          s2 += j2[k] * (sg[k] * sg[k]);
        }
        s2 += R(1e-8);
        const real sd = sqrt(s2);
        const real z = (x - dy[j]) / sd;
        real lp, dl_dloc, dl_ds;
        if (a.likelihood == 0) {
          lp = R(-0.5) * z * z - log(sd) - R(0.9189385332046727418);
          dl_dloc = z / sd;
          dl_ds = (z * z - R(1.0)) / sd;
        } else {
          const real az = fabs(z);
          lp = R(-0.4515827052894548647) - log(sd) - (az + ctx_log1p_exp_neg2a(az));
          const real tz = tanh(z);
          dl_dloc = tz / sd;
          dl_ds = (z * tz - R(1.0)) / sd;
        }
        acc[0] += lp;
#pragma unroll
        for (int d = 0; d < CTX_D_THETA; ++d) acc[1 + d] += dl_dloc * ds[d * CTX_S + j];
#pragma unroll
        for (int k = 0; k < CTX_S; ++k) acc[1 + CTX_D_THETA + k] += dl_ds * (j2[k] * sg[k] / sd);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < CTX_KR; ++k) {
    double v = (double)acc[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(CTX_FULL, v, d);
    if (lane == 0) a.out[(size_t)ch * CTX_KR + k] = (real)v;
  }
}
#endif  // CTX_TAN && !CTX_THETA_PTR


// =====================================================================================
// kernel K4: right-hand-side evaluation at N points (no solve)
// =====================================================================================
// Replaces vmap(stack) / vmap(MLP): Model._setup_rate_function (model.py:1027-1051),
// NeuralBase.rates (neuralbase.py:319-327), the surrogate branch mcmc.py:830-833.
struct ctx_rates_args {
  const real* t;       // [N] or [1]
  const real* y;       // [N,S]
  const real* theta;   // [N,P] or [P]
  const real* consts;  // [N,C] or [C]
  const real* y0;      // [N,S] or [S] ("{state}_init" values) or null -> y
  real* out;           // [N,S]
  long long N;
  int stride_t, stride_theta, stride_consts, stride_y0;
};

extern "C" __global__ void __launch_bounds__(128) ctx_rates(const ctx_rates_args a) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  real y[CTX_S], y0[CTX_S], dy[CTX_S], c[CTX_C > 0 ? CTX_C : 1];
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) y[i] = a.y[(size_t)n * CTX_S + i];
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) y0[i] = a.y0 ? a.y0[(size_t)n * a.stride_y0 + i] : y[i];
#pragma unroll
  for (int i = 0; i < CTX_C; ++i) c[i] = a.consts[(size_t)n * a.stride_consts + i];
#if CTX_THETA_PTR
  const real* th = a.theta + (size_t)n * a.stride_theta;
#else
  real th[CTX_P > 0 ? CTX_P : 1];
#pragma unroll
  for (int i = 0; i < CTX_P; ++i) th[i] = a.theta[(size_t)n * a.stride_theta + i];
#endif
  ctx_rhs(a.t[(size_t)n * a.stride_t], y, th, c, y0, dy);
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) a.out[(size_t)n * CTX_S + i] = dy[i];
}


// =====================================================================================
// kernel K1w: one WARP per trajectory (lane = state) for models whose state does not fit the
// registers of one thread
// =====================================================================================
// Same algorithm as ctx_solve_v0 / v1 (simulation.py:259-273 under vmap :346-355), laid out for
// large S: lane l owns states l, l+32, ... (CTXW_NSS per lane) with their 7 stage slopes in
// registers; the vector field is the generated lane-parallel form (tools/codegen.py
// generate_warp_form): stage state -> shared memory, every lane evaluates its terms (ODE
// right-hand sides / reaction rates, one structural class per slot so the instruction stream is
// shared), terms -> shared memory, every lane sums its ELL row of the stoichiometric matrix.
// The error norm is a warp-shuffle reduction, so accept / reject, the step size and the number
// of requested times covered by a step are warp-uniform: no divergence at all.  Saved points
// are written by all lanes at once: S contiguous values per requested time.
// Persistent warps claim trajectories from a global counter.
#if defined(CTX_HAVE_WARP) && !CTX_TAN
#define CTXW_NY (CTXW_NSS * 32)
#define CTXW_NRS (CTXW_NTS * 32 + 32)
#ifndef CTXW_MIN_BLOCKS
#define CTXW_MIN_BLOCKS 1
#endif

struct ctxw_ctx {
  int it[CTXW_NI];
  real rt[CTXW_NR];
  real pre[CTXW_NPRE];
  real* sy;
  real* sr;
  int lane;
};

__device__ __forceinline__ void ctxw_field(const ctxw_ctx& w, const real t, const real* yv,
                                           real* dy) {
#pragma unroll
  for (int u = 0; u < CTXW_NSS; ++u) w.sy[u * 32 + w.lane] = yv[u];
  __syncwarp();
  ctxw_terms(w.it, w.pre, t, w.sy, w.sr, w.lane);
  __syncwarp();
  ctxw_dy(w.it, w.rt, w.sr, dy);
}

extern "C" __global__ void __launch_bounds__(CTXW_WARPS * 32, CTXW_MIN_BLOCKS)
ctx_solve_w(const ctx_solve_args a) {
  __shared__ real sm_y[CTXW_WARPS][CTXW_NY];
  __shared__ real sm_r[CTXW_WARPS][CTXW_NRS];
  ctxw_ctx w;
  w.lane = threadIdx.x & 31;
  w.sy = sm_y[threadIdx.x >> 5];
  w.sr = sm_r[threadIdx.x >> 5];
  const int lane = w.lane;
#pragma unroll
  for (int k = 0; k < CTXW_NI; ++k) w.it[k] = ctxw_itab[k * 32 + lane];
#pragma unroll
  for (int k = 0; k < CTXW_NR; ++k) w.rt[k] = ctxw_rtab[k * 32 + lane];
  w.sr[CTXW_NTS * 32 + lane] = R(0.0);   // the column padded ELL entries point at
  const int T = a.T;
  bool valid[CTXW_NSS];
#pragma unroll
  for (int u = 0; u < CTXW_NSS; ++u) valid[u] = (u * 32 + lane) < CTX_S;

  while (true) {
    unsigned long long bb = 0;
    if (lane == 0) bb = atomicAdd(a.work_counter, 1ull);
    bb = __shfl_sync(CTX_FULL, bb, 0);
    if (bb >= (unsigned long long)a.B) break;
    const long long b = (long long)bb;
    CTX_ROWS(b, ro, ri);
    real y[CTXW_NSS], ynew[CTXW_NSS], yt[CTXW_NSS];
    real K[CTX_NST][CTXW_NSS];
#pragma unroll
    for (int u = 0; u < CTXW_NSS; ++u)
      y[u] = valid[u] ? a.y0[(size_t)ri * a.stride_y0 + u * 32 + lane] : R(0.0);
    ctxw_preload(w.it, a.theta + (size_t)ro * a.stride_theta,
                 a.consts + (size_t)ri * a.stride_consts, a.y0 + (size_t)ri * a.stride_y0, w.pre);
    const real* ts = a.ts + (size_t)ri * a.stride_ts;
    real* out = a.ys + (size_t)b * T * CTX_S;
    const real t_end = ts[T - 1];
    real tprev = a.t0_from_ts ? ts[0] : R(0.0);
    const real dt0_r = CTX_ROW_DT0(ts, T);
    real tnext = fmin(tprev + dt0_r, t_end);
    int idx = 0, nacc = 0, nrej = 0, status = 0;

    if (!(tprev < t_end)) {   // t0 == t1: every requested time is t0
      for (; idx < T; ++idx) {
#pragma unroll
        for (int u = 0; u < CTXW_NSS; ++u)
          if (valid[u]) out[(size_t)idx * CTX_S + u * 32 + lane] = y[u];
      }
    }
    ctxw_field(w, tprev, y, K[0]);

    for (int n = 0; tprev < t_end; ++n) {
      if (n >= a.max_steps) { status = CTX_ST_MAX; break; }
      const real dt = tnext - tprev;
      // ---- stages (same tableau macros as ctx_rk_stages; K holds f_i)
#if CTX_SOLVER == 0 || CTX_SOLVER == 1
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u) yt[u] = y[u] + dt * (A21 * K[0][u]);
      ctxw_field(w, tprev + C2 * dt, yt, K[1]);
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u) yt[u] = y[u] + dt * (A31 * K[0][u] + A32 * K[1][u]);
      ctxw_field(w, tprev + C3 * dt, yt, K[2]);
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u)
        yt[u] = y[u] + dt * (A41 * K[0][u] + A42 * K[1][u] + A43 * K[2][u]);
      ctxw_field(w, tprev + C4 * dt, yt, K[3]);
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u)
        yt[u] = y[u] + dt * (A51 * K[0][u] + A52 * K[1][u] + A53 * K[2][u] + A54 * K[3][u]);
      ctxw_field(w, tprev + C5 * dt, yt, K[4]);
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u)
        yt[u] = y[u] + dt * (A61 * K[0][u] + A62 * K[1][u] + A63 * K[2][u] + A64 * K[3][u] +
                             A65 * K[4][u]);
      ctxw_field(w, tprev + dt, yt, K[5]);
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u)
        ynew[u] = y[u] + dt * (A71 * K[0][u] + A72 * K[1][u] + A73 * K[2][u] + A74 * K[3][u] +
                               A75 * K[4][u] + A76 * K[5][u]);
      ctxw_field(w, tprev + dt, ynew, K[6]);
#elif CTX_SOLVER == 2
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u) ynew[u] = y[u] + dt * K[0][u];
#else
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u) yt[u] = y[u] + dt * K[0][u];
      ctxw_field(w, tprev + dt, yt, K[1]);
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u) ynew[u] = y[u] + dt * (R(0.5) * K[0][u] + R(0.5) * K[1][u]);
#endif
      bool keep = true;
      real dt_next = dt0_r;
#if CTX_ADAPTIVE
      {
        bool my_nan = false;
#pragma unroll
        for (int u = 0; u < CTXW_NSS; ++u) my_nan |= valid[u] && (ynew[u] != ynew[u]);
        const bool has_nan = __any_sync(CTX_FULL, my_nan);
        real ssq = R(0.0);
#pragma unroll
        for (int u = 0; u < CTXW_NSS; ++u) {
          real e = dt * (E1 * K[0][u] + E2 * K[1][u] + E3 * K[2][u] + E4 * K[3][u] +
                         E5 * K[4][u] + E6 * K[5][u] + E7 * K[6][u]);
          if (e != e) e = CTX_INF;
          const real y1 = has_nan ? y[u] : ynew[u];
          const real sc = a.atol + fmax(fabs(y[u]), fabs(y1)) * a.rtol;
          const real r = e / sc;
          if (valid[u]) ssq += r * r;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ssq += __shfl_xor_sync(CTX_FULL, ssq, d);
        ssq = __shfl_sync(CTX_FULL, ssq, 0);
        const real scaled = sqrt(ssq / (real)CTX_S);
        keep = scaled < R(1.0);
        const real inv = CTX_IDIV(R(1.0), scaled);
        real factor = R(0.9) * ctx_pow02(inv);
        factor = fmin(fmax(factor, keep ? R(1.0) : R(0.2)), R(10.0));
        if (factor != factor) factor = inv;
        dt_next = dt * factor;
      }
#endif
      if (keep) {
        // ---- dense output: polynomial in theta of every owned state, evaluated for the
        // requested times this step covers (lanes prefetch 32 times, then broadcast them)
        real cf[CTXW_NSS][CTX_DEG + 1];
#pragma unroll
        for (int u = 0; u < CTXW_NSS; ++u) {
#if CTX_SOLVER == 0
          const real d2 = K[1][u] - K[0][u], d3 = K[2][u] - K[0][u], d4 = K[3][u] - K[0][u],
                     d5 = K[4][u] - K[0][u], d6 = K[5][u] - K[0][u], d7 = K[6][u] - K[0][u];
          cf[u][0] = y[u];
          cf[u][1] = dt * K[0][u];
          cf[u][2] = dt * (TB22 * d2 + TB32 * d3 + TB42 * d4 + TB52 * d5 + TB62 * d6 + TB72 * d7);
          cf[u][3] = dt * (TB23 * d2 + TB33 * d3 + TB43 * d4 + TB53 * d5 + TB63 * d6 + TB73 * d7);
          cf[u][4] = dt * (TB24 * d2 + TB34 * d3 + TB44 * d4 + TB54 * d5 + TB64 * d6 + TB74 * d7);
#elif CTX_SOLVER == 1
          const real ymid = y[u] + dt * (M1 * K[0][u] + M2 * K[1][u] + M3 * K[2][u] + M4 * K[3][u] +
                                         M5 * K[4][u] + M6 * K[5][u] + M7 * K[6][u]);
          const real f0 = dt * K[0][u], f1 = dt * K[6][u];
          cf[u][0] = y[u];
          cf[u][1] = f0;
          cf[u][2] = f1 - R(4.0) * f0 - R(11.0) * y[u] - R(5.0) * ynew[u] + R(16.0) * ymid;
          cf[u][3] = R(5.0) * f0 - R(3.0) * f1 + R(18.0) * y[u] + R(14.0) * ynew[u] - R(32.0) * ymid;
          cf[u][4] = R(2.0) * (f1 - f0) - R(8.0) * (ynew[u] + y[u]) + R(16.0) * ymid;
#else
          cf[u][0] = y[u];
          cf[u][1] = ynew[u] - y[u];
#endif
        }
        const real inv_dt = (dt == R(0.0)) ? R(0.0) : R(1.0) / dt;
        while (idx < T) {
          const int j = idx + lane;
          const real tq = (j < T) ? ts[j] : R(0.0);
          const unsigned m = __ballot_sync(CTX_FULL, (j < T) && (tq <= tnext));
          const int cnt = (m == CTX_FULL) ? 32 : __ffs((int)~m) - 1;
          for (int q = 0; q < cnt; ++q) {
            const real th_ = (__shfl_sync(CTX_FULL, tq, q) - tprev) * inv_dt;
#pragma unroll
            for (int u = 0; u < CTXW_NSS; ++u) {
              real v = cf[u][CTX_DEG];
#pragma unroll
              for (int mm = CTX_DEG - 1; mm >= 0; --mm) v = v * th_ + cf[u][mm];
              if (valid[u]) out[(size_t)(idx + q) * CTX_S + u * 32 + lane] = v;
            }
          }
          idx += cnt;
          if (cnt < 32) break;
        }
#pragma unroll
        for (int u = 0; u < CTXW_NSS; ++u) y[u] = ynew[u];
#if CTX_SOLVER == 0 || CTX_SOLVER == 1
#pragma unroll
        for (int u = 0; u < CTXW_NSS; ++u) K[0][u] = K[6][u];
#endif
        tprev = tnext;
        ++nacc;
      } else {
        ++nrej;
      }
#if !(CTX_SOLVER == 0 || CTX_SOLVER == 1)
      if (keep) ctxw_field(w, tprev, y, K[0]);
#endif
      tprev = fmin(tprev, t_end);
      real tn = tprev + dt_next;
      if (tn > t_end - CTX_TOL_END) tn = keep ? t_end : tprev + R(0.5) * (t_end - tprev);
      tnext = tn;
      if (tprev < t_end && !(fabs(tnext) < CTX_INF)) { status = CTX_ST_BAD; break; }
    }
    if (tprev < t_end && status == 0) status = CTX_ST_MAX;
    for (; idx < T; ++idx) {
#pragma unroll
      for (int u = 0; u < CTXW_NSS; ++u)
        if (valid[u]) out[(size_t)idx * CTX_S + u * 32 + lane] = CTX_INF;
    }
    if (lane == 0) {
      a.status[b] = status;
      a.n_accepted[b] = nacc;
      a.n_rejected[b] = nrej;
    }
  }
}
#endif  // CTX_HAVE_WARP

#ifdef CTX_MLP_TC
#include "ctx_mlp_tc.cuh"
#endif
#ifdef CTX_MLP_ADJ
#include "ctx_mlp_adjoint.cuh"
#endif
