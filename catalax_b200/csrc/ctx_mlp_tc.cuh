// ctx_mlp_tc.cuh -- K3: batched MLP vector field on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Included at the end of ctx_integrator.cuh when the generated field defines CTX_MLP_TC
// (catalax_b200/tools/codegen_mlp.py).  Replaces, for f32 neural models, the XLA computation of
//   catalax/neural/neuralode.py:49 / rateflow.py:111 / universalode.py:130
//        vmap(diffeqsolve(ODETerm(field), solver, t0, t1, dt0, y0, PIDController, SaveAt(ts)))
//   with field = MLP (catalax/neural/mlp.py:71-84) [+ stoichiometry / gate / mechanistic term].
//
// One CTA = 512 threads = two independent tiles of 128 trajectories.  A tile has 8 warps: warps
// 0-3 ("primary", thread r <-> trajectory r <-> TMEM lane r) own the integrator state; warps 4-7
// ("helpers") only share the epilogue: warp q and warp q+4 may both read TMEM lanes 32q..32q+31,
// so each half handles half of the hidden units (twice the warps to hide MUFU/FMA latency).  Every RK stage evaluates the MLP for the whole tile as dense
// contractions  D[128 x N_l] = A[128 x K_l] . W_l[N_l x K_l]^T :
//   * A (activations) is written by the owning threads into shared memory in the UMMA
//     K-major / no-swizzle canonical layout (8x16B core matrices), W_l is staged there once;
//   * one elected thread per tile issues tcgen05.mma.kind::tf32 (M=128, N=N_l, K=8 per
//     instruction), accumulator in TMEM, completion via tcgen05.commit -> mbarrier;
//   * the epilogue reads the thread's accumulator row with tcgen05.ld (32x32b), adds the bias,
//     applies the activation and writes the next layer's A tile.
// Precision: the reference computes the MLP in f32.  TF32 alone (10-bit mantissa) cannot meet an
// f32 bound, so every product is split (3xTF32): a = a_hi + a_lo, w = w_hi + w_lo and
// D += a_hi.w_hi + a_hi.w_lo + a_lo.w_hi; accumulation is f32 in TMEM, bias/activation in f32.
// While one tile waits for its MMAs the other tile's threads run their epilogue, so the tensor
// pipe and the FP32/MUFU pipes overlap inside one SM.
//
// Lanes are refilled from a global counter as in ctx_solve_v1 (all 128 lanes of a tile attempt a
// step in every iteration); the first-stage slope is re-evaluated every iteration (7 collective
// evaluations per step) so that freshly claimed lanes need no special case.
#pragma once

#if CTX_F64
#error "the tensor-core MLP kernel is f32 only"
#endif

#define TC_ROWS 128
#define TC_THREADS 512
#define TC_TILE_THREADS 256
#define TC_HALF_W (MLP_W / 2)                         // hidden units handled per half
#define TC_NL (MLP_DEPTH + 1)
#define TC_KMAX (MLP_W > MLP_IN_PAD ? MLP_W : MLP_IN_PAD)
#define TC_A_BYTES (TC_ROWS * TC_KMAX * 4)          // one A tile (hi or lo)
#define TC_KBLK_STRIDE (TC_ROWS / 8 * 128)            // bytes between 4-element k blocks of A

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ unsigned tc_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float tc_tf32(const float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// K-major, no swizzle: start address, LBO = byte stride between core matrices along K,
// SBO = byte stride between core matrices along M/N, descriptor version 1 (sm_100).
__device__ __forceinline__ unsigned long long tc_desc(const unsigned saddr, const unsigned lbo,
                                                      const unsigned sbo) {
  return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | ((unsigned long long)(lbo >> 4) << 16) |
         ((unsigned long long)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::tf32 instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ constexpr unsigned tc_idesc(const int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(TC_ROWS >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(const unsigned d_tmem, const unsigned long long a,
                                       const unsigned long long b, const unsigned idesc,
                                       const unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(const unsigned mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(const unsigned mbar, const unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_bar(const int id) {  // the 128 threads of one tile
  asm volatile("bar.sync %0, 256;" :: "r"(id) : "memory");
}
__device__ __forceinline__ bool tc_tile_or(const int id, const bool pred) {
  unsigned out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\t"
      "bar.red.or.pred q, %2, 256, p;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
      : "=r"(out) : "r"((unsigned)pred), "r"(id) : "memory");
  return out != 0;
}
#define TC_LD16(v, off, taddr)                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                        \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                \
               : "=r"(v[off + 0]), "=r"(v[off + 1]), "=r"(v[off + 2]), "=r"(v[off + 3]),        \
                 "=r"(v[off + 4]), "=r"(v[off + 5]), "=r"(v[off + 6]), "=r"(v[off + 7]),        \
                 "=r"(v[off + 8]), "=r"(v[off + 9]), "=r"(v[off + 10]), "=r"(v[off + 11]),      \
                 "=r"(v[off + 12]), "=r"(v[off + 13]), "=r"(v[off + 14]), "=r"(v[off + 15])     \
               : "r"((taddr) + (off)))

// Activations of the tensor-core form.  The epilogue, not the MMAs, bounds this kernel (ncu:
// libdevice tanhf = 60 % of all instructions), and the operands already carry the 3xTF32
// splitting error (~5e-7 relative), so the transcendental part is evaluated with the SFU
// approximations (ex2 / lg2 / rcp .approx: <= 2 ulp each): absolute error <= ~3e-7 per
// activation, far inside the f32 bound of the solve (rtol 1e-3) and of the same order as the
// splitting error.  MLP_FAST_ACT=0 restores the libdevice functions.
#ifndef MLP_FAST_ACT
#define MLP_FAST_ACT 1
#endif
__device__ __forceinline__ float tc_ex2(const float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tc_lg2(const float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tc_rcp(const float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tc_act(const float x, const int kind) {
#if MLP_FAST_ACT
  switch (kind) {
    case 0:   // tanh = 1 - 2 / (exp(2x) + 1): saturates correctly (ex2 -> inf | 0)
      return fmaf(-2.0f, tc_rcp(tc_ex2(x * 2.8853900817779268f) + 1.0f), 1.0f);
    case 1:   // softplus = max(x, 0) + log(1 + exp(-|x|))
      return fmaf(0.69314718055994531f, tc_lg2(1.0f + tc_ex2(-1.4426950408889634f * fabsf(x))),
                  fmaxf(x, 0.0f));
    case 2: return x > 0.0f ? x : tc_ex2(1.4426950408889634f * x) - 1.0f;   // celu, alpha = 1
    case 3: return fmaxf(x, 0.0f);
    default: return x;
  }
#else
  switch (kind) {
    case 0: return tanhf(x);
    case 1: return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
    case 2: return x > 0.0f ? x : expm1f(x);
    case 3: return fmaxf(x, 0.0f);
    default: return x;
  }
#endif
}

struct tc_tile {
  unsigned a_hi, a_lo;    // shared-window addresses of this tile's A tiles
  unsigned mbar;          // mbarrier the tile's MMAs commit to
  unsigned tmem;          // TMEM address of the tile's accumulator (column offset)
  unsigned w_base;        // shared-window address of the staged UMMA weights
  const float* aux;       // biases + extras in shared memory
  int bar_id;             // named barrier of the tile's 128 threads
  int r;                  // row of this thread in the tile
  int half;               // 0 = primary (owns the trajectory), 1 = epilogue helper
  unsigned parity;
};

// write 4 consecutive K elements [k, k+4) of this thread's row into A_hi / A_lo
__device__ __forceinline__ void tc_store4(const tc_tile& T, const int k, const float v0,
                                          const float v1, const float v2, const float v3) {
  const unsigned off = (unsigned)(k >> 2) * TC_KBLK_STRIDE + (unsigned)(T.r >> 3) * 128u + (unsigned)(T.r & 7) * 16u;
  const float h0 = tc_tf32(v0), h1 = tc_tf32(v1), h2 = tc_tf32(v2), h3 = tc_tf32(v3);
  const float l0 = tc_tf32(v0 - h0), l1 = tc_tf32(v1 - h1), l2 = tc_tf32(v2 - h2), l3 = tc_tf32(v3 - h3);
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(T.a_hi + off), "f"(h0), "f"(h1), "f"(h2), "f"(h3) : "memory");
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(T.a_lo + off), "f"(l0), "f"(l1), "f"(l2), "f"(l3) : "memory");
}

// One collective evaluation of the field for the 128 trajectories of a tile.
// Every thread passes its own (t, y); all 128 threads must call it together.
__device__ __forceinline__ void tc_field(tc_tile& T, const float t, const float* y, const float* y0,
                                         float* dy) {
  // ---- layer-0 input row: [y, t/max_time, 0...]
  if (T.half == 0) {
    float x[MLP_IN_PAD];
#pragma unroll
    for (int i = 0; i < MLP_IN_PAD; ++i) x[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < MLP_S; ++i) x[i] = y[i];
    x[MLP_S] = t * T.aux[MLP_AUX_EXTRA];
#pragma unroll
    for (int k = 0; k < MLP_IN_PAD; k += 4) tc_store4(T, k, x[k], x[k + 1], x[k + 2], x[k + 3]);
  }
  unsigned w_off = 0;   // byte offset of layer l's weights (hi block, then lo block)
  int bias_off = 0;
#pragma unroll
  for (int l = 0; l < TC_NL; ++l) {
    const int K = (l == 0) ? MLP_IN_PAD : MLP_W;
    const int N = (l == TC_NL - 1) ? MLP_OUT_PAD : MLP_W;
    // A tile written with generic stores -> make it visible to the async proxy, then sync the tile
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tc_bar(T.bar_id);
    if (T.r == 0 && T.half == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned lbo_b = (unsigned)N * 16u;        // N/8 core matrices of 128 B per k block
      const unsigned w_hi = T.w_base + w_off, w_lo = w_hi + (unsigned)(N * K * 4);
      const unsigned idesc = tc_idesc(N);
#pragma unroll
      for (int ks = 0; ks < K / 8; ++ks) {
        const unsigned long long a_hi = tc_desc(T.a_hi + ks * 2 * TC_KBLK_STRIDE, TC_KBLK_STRIDE, 128);
        const unsigned long long a_lo = tc_desc(T.a_lo + ks * 2 * TC_KBLK_STRIDE, TC_KBLK_STRIDE, 128);
        const unsigned long long b_hi = tc_desc(w_hi + ks * 2 * lbo_b, lbo_b, 128);
        const unsigned long long b_lo = tc_desc(w_lo + ks * 2 * lbo_b, lbo_b, 128);
        tc_mma(T.tmem, a_lo, b_hi, idesc, ks > 0 ? 1u : 0u);   // small terms first
        tc_mma(T.tmem, a_hi, b_lo, idesc, 1u);
        tc_mma(T.tmem, a_hi, b_hi, idesc, 1u);
      }
      tc_commit(T.mbar);
    }
    tc_mbar_wait(T.mbar, T.parity);
    T.parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: this thread's accumulator row (TMEM lane = 32*(warp%4) + lane)
    const unsigned taddr = T.tmem + ((unsigned)((T.r >> 5) * 32) << 16);
    if (l < TC_NL - 1) {
      // this half's hidden units [c, c + W/2)
      const int c = T.half * TC_HALF_W;
      unsigned v[TC_HALF_W];
#pragma unroll
      for (int c0 = 0; c0 < TC_HALF_W; c0 += 16) TC_LD16(v, c0, taddr + c);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const float* bias = T.aux + bias_off + c;
#pragma unroll
      for (int k = 0; k < TC_HALF_W; k += 4) {
        const float a0 = tc_act(__uint_as_float(v[k]) + bias[k], MLP_ACT);
        const float a1 = tc_act(__uint_as_float(v[k + 1]) + bias[k + 1], MLP_ACT);
        const float a2 = tc_act(__uint_as_float(v[k + 2]) + bias[k + 2], MLP_ACT);
        const float a3 = tc_act(__uint_as_float(v[k + 3]) + bias[k + 3], MLP_ACT);
        tc_store4(T, c + k, a0, a1, a2, a3);
      }
    } else if (T.half == 0) {
      unsigned v[MLP_OUT_PAD];
#pragma unroll
      for (int c0 = 0; c0 < MLP_OUT_PAD; c0 += 16) TC_LD16(v, c0, taddr);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float o[MLP_OUT];
#pragma unroll
      for (int j = 0; j < MLP_OUT; ++j)
        o[j] = tc_act(__uint_as_float(v[j]) + T.aux[bias_off + j], MLP_FINAL_ACT);
      ctx_mlp_finish(t, y, y0, T.aux + MLP_AUX_EXTRA, o, dy);
    }
    w_off += (unsigned)(2 * N * K * 4);
    bias_off += N;
  }
}

// theta = [aux (MLP_AUX_SIZE floats) | UMMA weight pack (MLP_WPACK_SIZE floats)]
extern "C" __global__ void __launch_bounds__(TC_THREADS, 1) ctx_solve_mlp_tc(const ctx_solve_args a) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  // layout: [weights][A_hi t0][A_lo t0][A_hi t1][A_lo t1][aux][mbar x2][tmem ptr]
  float* w_sm = reinterpret_cast<float*>(tc_smem);
  unsigned char* a_sm = tc_smem + ((MLP_WPACK_SIZE * 4 + 1023) / 1024) * 1024;
  float* aux_sm = reinterpret_cast<float*>(a_sm + 4 * TC_A_BYTES);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(aux_sm + ((MLP_AUX_SIZE + 3) / 4) * 4);
  unsigned* tmem_ptr = reinterpret_cast<unsigned*>(mbar + 2);

  const int tid = threadIdx.x;
  for (int i = tid; i < MLP_WPACK_SIZE; i += TC_THREADS) w_sm[i] = a.theta[MLP_AUX_SIZE + i];
  for (int i = tid; i < MLP_AUX_SIZE; i += TC_THREADS) aux_sm[i] = a.theta[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(&mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(&mbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(tc_smem_u32(tmem_ptr)), "n"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged weights -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const int wg = tid >> 8;                         // tile
  const int w_in = (tid >> 5) & 7;                 // warp within the tile
  tc_tile T;
  T.half = w_in >> 2;
  T.r = (w_in & 3) * 32 + (tid & 31);
  T.bar_id = 1 + wg;
  T.a_hi = tc_smem_u32(a_sm + (2 * wg) * TC_A_BYTES);
  T.a_lo = tc_smem_u32(a_sm + (2 * wg + 1) * TC_A_BYTES);
  T.mbar = tc_smem_u32(&mbar[wg]);
  T.tmem = *tmem_ptr + (unsigned)(wg * MLP_W);
  T.w_base = tc_smem_u32(w_sm);
  T.aux = aux_sm;
  T.parity = 0u;

  const int lane = tid & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int Tn = a.T;
  long long b = -1;
  float y[CTX_S], ynew[CTX_S], y0[CTX_S], yt[CTX_S];
  float K[CTX_NST][CTX_S];
  float tprev = 0.0f, tnext = 0.0f, t_end = 0.0f;
  int idx = 0, nacc = 0, nrej = 0;
  bool exhausted = false;
  const float* tsr = a.ts;
#pragma unroll
  for (int i = 0; i < CTX_S; ++i) y[i] = y0[i] = 1.0f;

  while (true) {
    // ---- refill free lanes (warp-aggregated claim; helper warps own no trajectory)
    const unsigned need = T.half == 0 ? __ballot_sync(CTX_FULL, b < 0) : 0u;
    if (need && !exhausted) {
      const int cnt = __popc(need);
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.work_counter, (unsigned long long)cnt);
      base = __shfl_sync(CTX_FULL, base, 0);
      if (base + cnt >= (unsigned long long)a.B) exhausted = true;
      if (b < 0) {
        const unsigned long long nb = base + __popc(need & lt_mask);
        if (nb < (unsigned long long)a.B) {
          b = (long long)nb;
          CTX_ROWS(b, ro, ri);
          (void)ro;
#pragma unroll
          for (int i = 0; i < CTX_S; ++i) y0[i] = y[i] = a.y0[(size_t)ri * a.stride_y0 + i];
          tsr = a.ts + (size_t)ri * a.stride_ts;
          t_end = tsr[Tn - 1];
          tprev = a.t0_from_ts ? tsr[0] : 0.0f;
          tnext = fminf(tprev + CTX_ROW_DT0(tsr, Tn), t_end);
          idx = 0; nacc = 0; nrej = 0;
          if (!(tprev < t_end)) {                    // t0 == t1: every requested time is t0
            for (; idx < Tn; ++idx)
#pragma unroll
              for (int i = 0; i < CTX_S; ++i) a.ys[((size_t)b * Tn + idx) * CTX_S + i] = y[i];
            a.status[b] = 0; a.n_accepted[b] = 0; a.n_rejected[b] = 0;
            b = -1;
          }
        }
      }
    }
    if (!tc_tile_or(T.bar_id, b >= 0)) break;        // tile-uniform: all 128 lanes are done

    // ---- one step attempt for every lane of the tile (idle lanes compute on stale data)
    const float dt = tnext - tprev;
    tc_field(T, tprev, y, y0, K[0]);
#if CTX_SOLVER == 0 || CTX_SOLVER == 1
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) yt[i] = y[i] + dt * (A21 * K[0][i]);
    tc_field(T, tprev + C2 * dt, yt, y0, K[1]);
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) yt[i] = y[i] + dt * (A31 * K[0][i] + A32 * K[1][i]);
    tc_field(T, tprev + C3 * dt, yt, y0, K[2]);
#pragma unroll
    for (int i = 0; i < CTX_S; ++i) yt[i] = y[i] + dt * (A41 * K[0][i] + A42 * K[1][i] + A43 * K[2][i]);
    tc_field(T, tprev + C4 * dt, yt, y0, K[3]);
#pragma unroll
    for (int i = 0; i < CTX_S; ++i)
      yt[i] = y[i] + dt * (A51 * K[0][i] + A52 * K[1][i] + A53 * K[2][i] + A54 * K[3][i]);
    tc_field(T, tprev + C5 * dt, yt, y0, K[4]);
#pragma unroll
    for (int i = 0; i < CTX_S; ++i)
      yt[i] = y[i] + dt * (A61 * K[0][i] + A62 * K[1][i] + A63 * K[2][i] + A64 * K[3][i] + A65 * K[4][i]);
    tc_field(T, tprev + dt, yt, y0, K[5]);
#pragma unroll
    for (int i = 0; i < CTX_S; ++i)
      ynew[i] = y[i] + dt * (A71 * K[0][i] + A72 * K[1][i] + A73 * K[2][i] + A74 * K[3][i] +
                             A75 * K[4][i] + A76 * K[5][i]);
    tc_field(T, tprev + dt, ynew, y0, K[6]);
#else
#error "the tensor-core MLP kernel supports Tsit5 and Dopri5"
#endif
    if (b >= 0) {
      const float scaled = ctx_scaled_error(dt, y, ynew, K, a.rtol, a.atol);
      const bool keep = scaled < 1.0f;
      const float inv = 1.0f / scaled;
      float factor = 0.9f * __powf(inv, 0.2f);
      factor = fminf(fmaxf(factor, keep ? 1.0f : 0.2f), 10.0f);
      float dt_next = dt * factor;
      if (scaled != scaled) dt_next = scaled;
      int status = 0;
      if (keep) {
        float out[CTX_S];
        while (idx < Tn) {
          const float tq = tsr[idx];
          if (!(tq <= tnext)) break;
          ctx_dense(tq, tprev, tnext, y, ynew, K, out);
#pragma unroll
          for (int i = 0; i < CTX_S; ++i) a.ys[((size_t)b * Tn + idx) * CTX_S + i] = out[i];
          ++idx;
        }
#pragma unroll
        for (int i = 0; i < CTX_S; ++i) y[i] = ynew[i];
        tprev = tnext;
        ++nacc;
      } else {
        ++nrej;
      }
      tprev = fminf(tprev, t_end);
      float tn = tprev + dt_next;
      if (tn > t_end - CTX_TOL_END) tn = keep ? t_end : tprev + 0.5f * (t_end - tprev);
      tnext = tn;
      const bool done = !(tprev < t_end);
      if (!done) {
        if (!(fabsf(tn) < CTX_INF)) status = CTX_ST_BAD;
        else if (nacc + nrej >= a.max_steps) status = CTX_ST_MAX;
      }
      if (done || status != 0) {
        for (; idx < Tn; ++idx)
#pragma unroll
          for (int i = 0; i < CTX_S; ++i) a.ys[((size_t)b * Tn + idx) * CTX_S + i] = CTX_INF;
        a.status[b] = status; a.n_accepted[b] = nacc; a.n_rejected[b] = nrej;
        b = -1;
        tprev = 0.0f; tnext = 0.0f; t_end = 0.0f;
      }
    }
  }
  __syncthreads();
  if (tid < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(*tmem_ptr), "n"(TC_TMEM_COLS) : "memory");
}
