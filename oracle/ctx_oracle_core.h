/* ctx_oracle_core.h -- plain-C restatement of the reference's batched ODE solve.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Included by a generated translation unit that
 * first defines `real`, ORACLE_S (states), ORACLE_N (states + tangents), ORACLE_NERR and the
 * model's `oracle_rhs()` (printed by sympy's own C printer, see oracle/c_oracle.py).
 *
 * Same algorithm, same operation order as oracle/diffrax_oracle.py (which cites the reference:
 * catalax/tools/simulation.py:236-275,346-355 and diffrax 0.7's Tsit5 / PIDController /
 * SaveAt(ts)); one trajectory per OpenMP task instead of numpy's masked lock-step, which is
 * what jax.vmap's semantics amount to (every lane has its own dt and accept decision).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifndef ORACLE_F32
#define ORACLE_F32 0
#endif
#if ORACLE_F32
#define RC(x) x##f
#define R_POW powf
#define R_SQRT sqrtf
#define R_FABS fabsf
#define R_FMAX fmaxf
#define R_FMIN fminf
#define TOL_END 1e-6f
#else
#define RC(x) x
#define R_POW pow
#define R_SQRT sqrt
#define R_FABS fabs
#define R_FMAX fmax
#define R_FMIN fmin
#define TOL_END 1e-10
#endif

static const real A_[7][6] = {
    {0},
    {RC(0.161)},
    {RC(-0.008480655492356988544426874250230774675121177393430391537369234245294192976164141156943),
     RC(0.3354806554923569885444268742502307746751211773934303915373692342452941929761641411569)},
    {RC(2.897153057105493432130432594192938764924887287701866490314866693455023795137503079289),
     RC(-6.359448489975074843148159912383825625952700647415626703305928850207288721235210244366),
     RC(4.362295432869581411017727318190886861027813359713760212991062156752264926097707165077)},
    {RC(5.325864828439256604428877920840511317836476253097040101202360397727981648835607691791),
     RC(-11.74888356406282787774717033978577296188744178259862899288666928009020615663593781589),
     RC(7.495539342889836208304604784564358155658679161518186721010132816213648793440552049753),
     RC(-0.09249506636175524925650207933207191611349983406029535244034750452930469056411389539635)},
    {RC(5.861455442946420028659251486982647890394337666164814434818157239052507339770711679748),
     RC(-12.92096931784710929170611868178335939541780751955743459166312250439928519268343184452),
     RC(8.159367898576158643180400794539253485181918321135053305748355423955009222648673734986),
     RC(-0.07158497328140099722453054252582973869127213147363544882721139659546372402303777878835),
     RC(-0.02826905039406838290900305721271224146717633626879770007617876201276764571291579142206)},
    {RC(0.09646076681806522951816731316512876333711995238157997181903319145764851595234062815396),
     RC(0.01),
     RC(0.4798896504144995747752495322905965199130404621990332488332634944254542060153074523509),
     RC(1.379008574103741893192274821856872770756462643091360525934940067397245698027561293331),
     RC(-3.290069515436080679901047585711363850115683290894936158531296799594813811049925401677),
     RC(2.324710524099773982415355918398765796109060233222962411944060046314465391054716027841)}};
static const real C_[7] = {RC(0.0), RC(0.161), RC(0.327), RC(0.9),
                           RC(0.9800255409045096857298102862870245954942137979563024768854764293221195950761080302604),
                           RC(1.0), RC(1.0)};
/* b_sol - b_hat, computed in double (oracle/tableaus.py TSIT5_E) */
static const real E_[7] = {RC(0.0017800110522257773), RC(0.0008164344596567463),
                           RC(-0.007880878010261994), RC(0.1447110071732629),
                           RC(-0.5823571654525552),   RC(0.45808210592918686),
                           RC(-0.015151515151515152)};

static void dense_weights(real t, real* b) {
  const real t2 = t * t;
  b[0] = RC(-1.0530884977290216) * t * (t - RC(1.3299890189751412)) *
         (t2 - RC(1.4364028541716351) * t + RC(0.7139816917074209));
  b[1] = RC(0.1017) * t2 * (t2 - RC(2.1966568338249754) * t + RC(1.2949852507374631));
  b[2] = RC(2.490627285651252793) * t2 * (t2 - RC(2.38535645472061657) * t + RC(1.57803468208092486));
  b[3] = RC(-16.54810288924490272) * (t - RC(1.21712927295533244)) * (t - RC(0.61620406037800089)) * t2;
  b[4] = RC(47.37952196281928122) * (t - RC(1.203071208372362603)) * (t - RC(0.658047292653547382)) * t2;
  b[5] = RC(-34.87065786149660974) * (t - RC(1.2)) * (t - RC(0.666666666666666667)) * t2;
  b[6] = RC(2.5) * (t - RC(1.0)) * (t - RC(0.6)) * t2;
}

/* One trajectory, Tsit5 + PID defaults + SaveAt(ts).  out: [T][ORACLE_N]. */
static void solve_one(const real* y0, const real* th, const real* c, const real* init_aug,
                      const real* ts, int T, real rtol, real atol, real dt0, int max_steps,
                      real* out, int32_t* status, int32_t* nacc, int32_t* nrej) {
  real y[ORACLE_N], y1[ORACLE_N], yi[ORACLE_N], f0[ORACLE_N], k[7][ORACLE_N], fi[ORACLE_N];
  const real inf = (real)INFINITY;
  for (int i = 0; i < ORACLE_S; ++i) y[i] = y0[i];
  for (int i = ORACLE_S; i < ORACLE_N; ++i) y[i] = init_aug ? init_aug[i - ORACLE_S] : RC(0.0);
  for (int j = 0; j < T * ORACLE_N; ++j) out[j] = inf;
  const real t_end = ts[T - 1];
  real tprev = RC(0.0), tnext = R_FMIN(tprev + dt0, t_end);
  int idx = 0, st = 0, na = 0, nr = 0, nsteps = 0;
  if (!(tprev < t_end)) {
    for (int j = 0; j < T; ++j)
      for (int i = 0; i < ORACLE_N; ++i) out[j * ORACLE_N + i] = y[i];
    idx = T;
  }
  oracle_rhs(tprev, y, th, c, y0, f0);
  while (tprev < t_end && st == 0) {
    if (nsteps >= max_steps) { st = 1; break; }
    const real dt = tnext - tprev;
    for (int s = 0; s < 7; ++s) {
      if (s == 0) {
        for (int i = 0; i < ORACLE_N; ++i) { fi[i] = f0[i]; }
      } else {
        for (int i = 0; i < ORACLE_N; ++i) {
          real acc = A_[s][0] * k[0][i];
          for (int j = 1; j < s; ++j) acc = acc + A_[s][j] * k[j][i];
          yi[i] = y[i] + acc;
        }
        oracle_rhs(tprev + C_[s] * dt, yi, th, c, y0, fi);
      }
      for (int i = 0; i < ORACLE_N; ++i) k[s][i] = fi[i] * dt;
    }
    for (int i = 0; i < ORACLE_N; ++i) y1[i] = yi[i]; /* SSAL: y1 = last stage input */
    /* error norm over the first ORACLE_NERR components */
    int has_nan = 0;
    for (int i = 0; i < ORACLE_NERR; ++i) has_nan |= (y1[i] != y1[i]);
    real ssq = RC(0.0);
    for (int i = 0; i < ORACLE_NERR; ++i) {
      real e = E_[0] * k[0][i];
      for (int j = 1; j < 7; ++j) e = e + E_[j] * k[j][i];
      if (e != e) e = inf;
      const real ys = has_nan ? y[i] : y1[i];
      const real sc = atol + R_FMAX(R_FABS(y[i]), R_FABS(ys)) * rtol;
      const real r = e / sc;
      ssq += r * r;
    }
    const real scaled = R_SQRT(ssq / (real)ORACLE_NERR);
    const int keep = scaled < RC(1.0);
    const real inv = RC(1.0) / scaled;
    real factor = RC(0.9) * R_POW(inv, RC(0.2));
    const real lo = keep ? RC(1.0) : RC(0.2);
    factor = (factor != factor) ? factor : R_FMIN(R_FMAX(factor, lo), RC(10.0));
    const real new_dt = dt * factor;
    if (keep) {
      while (idx < T && ts[idx] <= tnext) {
        const real den = tnext - tprev;
        const real th_ = (den == RC(0.0)) ? RC(0.0) : (ts[idx] - tprev) / den;
        real b[7];
        dense_weights(th_, b);
        for (int i = 0; i < ORACLE_N; ++i) {
          real acc = b[0] * k[0][i];
          for (int j = 1; j < 7; ++j) acc = acc + b[j] * k[j][i];
          out[idx * ORACLE_N + i] = y[i] + acc;
        }
        ++idx;
      }
      for (int i = 0; i < ORACLE_N; ++i) { y[i] = y1[i]; f0[i] = fi[i]; }
      ++na;
    } else {
      ++nr;
    }
    ++nsteps;
    real ntp = keep ? tnext : tprev;
    ntp = R_FMIN(ntp, t_end);
    real ntn = ntp + new_dt;
    if (ntn > t_end - TOL_END) ntn = keep ? t_end : ntp + RC(0.5) * (t_end - ntp);
    tprev = ntp; tnext = ntn;
    if (!isfinite(ntn) && ntp < t_end) st = 2;
  }
  *status = st; *nacc = na; *nrej = nr;
}

/* Batched entry point: strides are elements per row, 0 = shared across the batch. */
void oracle_solve(int64_t B, int T, const real* y0, int64_t s_y0, const real* th, int64_t s_th,
                  const real* c, int64_t s_c, const real* ts, int64_t s_ts, const real* init_aug,
                  double rtol, double atol, double dt0, int max_steps, real* ys, int32_t* status,
                  int32_t* nacc, int32_t* nrej, int n_threads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads)
#endif
  for (int64_t b = 0; b < B; ++b) {
    solve_one(y0 + b * s_y0, th + b * s_th, c + b * s_c, init_aug, ts + b * s_ts, T, (real)rtol,
              (real)atol, (real)dt0, max_steps, ys + (size_t)b * T * ORACLE_N, status + b, nacc + b,
              nrej + b);
  }
}
