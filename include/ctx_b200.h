/* ctx_b200.h -- C ABI of the B200-native Catalax solve engine (libctxb200.so).
 *
 * The reference (JR-1991/Catalax) has NO native/FFI interface: its seam for this path is the
 * Python callable built by catalax/tools/simulation.py:160-196 (`Simulation._prepare_func`),
 *     f(y0, parameters, constants, time) -> ys            [B,T,S]
 * stored as Model._sim_func (catalax/model/model.py:786) and called from
 * model.py:699 (simulate), mcmc/mcmc.py:448 (NUTS log-density), mcmc/loo.py:181,
 * mcmc/predictive.py:151.  Every entry point below names the reference interface it replaces;
 * INTEGRATION.md shows the reference-side binding (ctypes / jax.ffi) a maintainer would add.
 *
 * Conventions
 *  - plain C types only; all array arguments of the *device* entry points are DEVICE pointers
 *    owned by the caller; outputs are caller-allocated; nothing is aliased or retained;
 *  - `stream` is a cudaStream_t passed as void*; calls are stream-ordered and never
 *    synchronise the device (the *_host variants synchronise the stream they are given);
 *  - dtype (f32 | f64) is fixed per compiled model, like jax's x64 switch
 *    (catalax/__init__.py:61-70); `real` below means that type;
 *  - every function returns 0 on success, a negative CTX_E* code otherwise, and
 *    ctx_last_error() returns the message of the last failure on the calling thread.
 */
#ifndef CTX_B200_H_
#define CTX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTX_ABI_VERSION 4

enum { CTX_F32 = 0, CTX_F64 = 1 };
/* solver ids: simconfig.py:16 (`solver: Any = Tsit5`), simulation.py:23-26 (non-adaptive list) */
enum { CTX_TSIT5 = 0, CTX_DOPRI5 = 1, CTX_EULER = 2, CTX_HEUN = 3 };
/* per-trajectory status written by the solve kernels (diffrax RESULTS, throw=False path,
 * simulation.py:255-257,271-272) */
enum { CTX_TRAJ_OK = 0, CTX_TRAJ_MAX_STEPS = 1, CTX_TRAJ_BAD_STEP = 2 };
/* likelihood ids: mcmc.py:59 (SoftLaplace default), Normal used by tests/docs */
enum { CTX_LIK_NORMAL = 0, CTX_LIK_SOFTLAPLACE = 1 };

enum {
  CTX_OK = 0,
  CTX_E_INVALID = -1,  /* bad argument */
  CTX_E_NVRTC = -2,    /* run-time compilation failed (log in ctx_last_error) */
  CTX_E_CUDA = -3,     /* CUDA runtime error */
  CTX_E_NOGPU = -4,    /* no CUDA device: there is NO CPU fallback */
  CTX_E_WORKSPACE = -5 /* workspace too small */
};

typedef struct ctx_model ctx_model;

/* Sizes baked into the generated vector-field source; checked against its CTX_* macros. */
typedef struct ctx_model_desc {
  int32_t n_states;     /* S: Model.get_state_order() (model.py:906-926) */
  int32_t n_params;     /* P: Model.get_parameter_order() (model.py:882-884) */
  int32_t n_consts;     /* C: Model.get_constants_order() (model.py:984-994) */
  int32_t n_dirs;       /* D: sensitivity directions generated (0 = none) */
  int32_t n_theta_dirs; /* leading D_theta directions are parameters, the rest initial values */
  int32_t dtype;        /* CTX_F32 | CTX_F64 */
} ctx_model_desc;

/* Frozen attributes of one solve = SimulationConfig (simconfig.py:8-20) + in_axes
 * (inaxes.py:5-9; simulation.py:346-355). */
typedef struct ctx_solve_cfg {
  int32_t solver;      /* CTX_TSIT5 ... */
  int32_t in_axes[4];  /* (y0, parameters, constants, time): 0 = batched on axis 0, -1 = None */
  int64_t batch;       /* B */
  int32_t n_times;     /* T = len(time) */
  int32_t max_steps;   /* simconfig.py:19 */
  double rtol, atol, dt0;
  int32_t kernel;      /* 0 = auto, 1 = static one-thread-per-trajectory, 2 = persistent refill,
                          3 = tcgen05 MLP tiles, 4 = one warp per trajectory (lane = state) */
  int32_t t0_from_ts;  /* 0: integrate from t0 = 0 (simulation.py:262); 1: from time[0], as the
                          neural solves do (neuralode.py:52, rateflow.py:114) */
  int32_t inner;       /* 0: one row index b for all operands.  M > 0: two-level batch
                          b = d*M + m -- parameters are indexed by d, y0 / constants / time by m:
                          vmap(per_draw)(thetas) over sim_func(y0s, theta, constants, times)
                          (predictive.py:150-154, loo.py:180-193); ys is [D, M, T, S] */
  int32_t reserved;    /* flags; bit 1 (value 2): run the scheduling probe (below) when `order` is NULL */
  const void* order;   /* optional DEVICE pointer to an int32 permutation of [0, batch): the persistent
                          kernels hand out rows in this order.  A scheduling hint only -- results do
                          not depend on it -- meant for "most expensive first" from the step counts of
                          a previous solve of similar inputs (a sampler's previous evaluation): one
                          launch is bound by the critical path of its longest trajectories, which then
                          start first.  MUST be a permutation (a missing row is never solved).
                          NULL = index order -- or, with flag bit 1 set, the order found by the
                          library's scheduling probe (a solve of the same batch at 100 x rtol capped at
                          32 step attempts; rows that hit the cap start first; csrc/ctx_sched.cu;
                          opt-in: it pays only where looser tolerances keep the long rows long).
                          Ignored by the static kernels and the host-buffer calls. */
} ctx_solve_cfg;

/* ---- model lifetime -------------------------------------------------------------------
 * Replaces building the jitted closure: Simulation._prepare_func (simulation.py:160-196) and
 * the sympy->JAX tracing of symbolicmodule.py:320-383.  `field_src` is the CUDA text emitted
 * by catalax_b200/tools/codegen.py.  Kernels are NVRTC-compiled for sm_100a lazily per
 * (solver, tangent) variant and cached (in process + on disk, CTX_B200_CACHE). */
int ctx_model_compile(const char* field_src, const ctx_model_desc* desc, ctx_model** out);
void ctx_model_free(ctx_model* m);
/* Compile one kernel variant WITHOUT a GPU and return its cubin (for build checks, cuobjdump).
 * The pointer stays valid until ctx_model_free. */
int ctx_model_cubin(ctx_model* m, int32_t solver, int32_t with_tangents, const void** cubin,
                    size_t* nbytes);

/* ---- K1: batched solve ------------------------------------------------------------------
 * Replaces `self._sim_func(y0, parameters, constants, saveat)` (model.py:699) =
 * vmap(diffeqsolve(...)) (simulation.py:259-273,348).
 * y0 [B,S]|[S], theta [B,P]|[P], consts [B,C]|[C], ts [B,T]|[T]  ->  ys [B,T,S];
 * status / n_accepted / n_rejected [B] (int32, may be NULL). */
size_t ctx_workspace_bytes(const ctx_model* m, const ctx_solve_cfg* cfg);
int ctx_solve(ctx_model* m, const ctx_solve_cfg* cfg, const void* y0, const void* theta,
              const void* consts, const void* ts, void* ys, int32_t* status,
              int32_t* n_accepted, int32_t* n_rejected, void* workspace, size_t ws_bytes,
              void* stream);

/* ---- K1 + tangents --------------------------------------------------------------------
 * Replaces vmap(jax.jacobian(solve, argnums=k)) (simulation.py:279-323): additionally writes
 * sens [B,T,S,D] = d ys / d(direction), directions = parameters (D_theta) then initial
 * values (D_y0), in the layout jax.jacobian produces ([T,S,arg], simulation.py:315). */
int ctx_solve_sens(ctx_model* m, const ctx_solve_cfg* cfg, const void* y0, const void* theta,
                   const void* consts, const void* ts, void* ys, void* sens, int32_t* status,
                   int32_t* n_accepted, int32_t* n_rejected, void* workspace, size_t ws_bytes,
                   void* stream);

/* ---- K2: fused log-likelihood + gradient ---------------------------------------------------
 * Replaces value_and_grad of the numpyro model body, catalax/mcmc/mcmc.py:405-491:
 *   states = sim_func(y0s, theta, constants, times)            (:448, in_axes=(0,None,0,0) :860)
 *   log L  = sum over mask of log likelihood(data; loc = states[..., observables],
 *                                            scale = sqrt(sigma_y**2))         (:461-491)
 * for `n_chains` independent (theta, sigma_y) pairs at once over the same M series.
 * The model must have been generated with parameter directions (and, to get d/dy0, initial
 * value directions).  Shapes (device pointers, model dtype):
 *   y0s [M,S]  theta [CH,P]  consts [M,C]  ts [M,T]  data [M,T,S]  mask [M,T,S] uint8 or NULL
 *   (NULL = isfinite(data), mcmc.py:566)  sigma [CH,S]
 *   out [CH, 1 + D_theta + S] = (log L, dlogL/dtheta, dlogL/dsigma_y) summed over the M series
 *   partial [CH*M, 1 + D_theta + S + D_y0]: per-(chain, series) rows; its last D_y0 columns are
 *   dlogL/dy0 of that series (init_estimator.py:47-58).  partial may be NULL (kept in workspace).
 * Series-sharded multi-GPU runs all-reduce `out` (sum) across ranks; chain-sharded runs need no
 * collective.  Prior terms (priors.py, error.py) are scalar and stay on the host side. */
typedef struct ctx_loglik_cfg {
  int32_t solver;
  int32_t likelihood; /* CTX_LIK_NORMAL | CTX_LIK_SOFTLAPLACE */
  int64_t n_chains;   /* CH */
  int32_t n_series;   /* M */
  int32_t n_times;    /* T */
  int32_t max_steps;
  int32_t kernel;     /* 0 = automatic (cooperative), 1 = per-lane likelihood terms, 2 = cooperative */
  double rtol, atol, dt0;
  const void* order;  /* optional DEVICE int32 permutation of [0, n_chains * n_series): scheduling hint
                         as in ctx_solve_cfg (trajectory u = chain * n_series + series) */
} ctx_loglik_cfg;
size_t ctx_loglik_workspace_bytes(const ctx_model* m, const ctx_loglik_cfg* cfg);
int ctx_loglik_grad(ctx_model* m, const ctx_loglik_cfg* cfg, const void* y0s, const void* theta,
                    const void* consts, const void* ts, const void* data, const uint8_t* mask,
                    const void* sigma, void* out, void* partial, int32_t* status, void* workspace,
                    size_t ws_bytes, void* stream);

/* ---- fused point-wise log-likelihood of posterior draws ------------------------------------------
 * Replaces `jax.vmap(per_draw)(thetas, sigma_y)` of catalax/mcmc/loo.py:180-193
 * (`states = sim_func(y0s, theta, constants, times); likelihood(states[..., observables],
 * scale).log_prob(data_safe)`): the solve of ctx_solve -- normally the two-level batch
 * cfg->inner = M, theta [D,P] per draw, y0 / consts / ts per series -- whose dense-output pass
 * writes  ll[d, m, t, o] = log p(data[m,t,o] | loc = y_o(t), scale = |sigma[d,o]|)  instead of the
 * states: the [D,M,T,S] trajectory block never exists.  data [M,T,S] must be finite (0 where
 * masked: `data_safe`; the caller applies the mask to ll), sigma [D,S]; without inner, data / sigma
 * rows follow the trajectory index.  Same workspace and status outputs as ctx_solve. */
int ctx_pointwise_loglik(ctx_model* m, const ctx_solve_cfg* cfg, int32_t likelihood, const void* y0,
                         const void* theta, const void* consts, const void* ts, const void* data,
                         const void* sigma, void* ll, int32_t* status, int32_t* n_accepted,
                         int32_t* n_rejected, void* workspace, size_t ws_bytes, void* stream);

/* ---- K4: right-hand side only ---------------------------------------------------------------
 * Replaces vmap(stack)(t, y, (theta, consts)) -- Model._setup_rate_function (model.py:1027-1051),
 * NeuralBase.rates (neuralbase.py:319-327), surrogate rate matching (mcmc.py:830-833).
 * t [N]|[1], y [N,S], theta [N,P]|[P], consts [N,C]|[C], y0 [N,S]|[S]|NULL -> out [N,S];
 * a stride of 0 elements means "shared across the N points". */
int ctx_rates(ctx_model* m, int64_t n, const void* t, int32_t stride_t, const void* y,
              const void* theta, int32_t stride_theta, const void* consts, int32_t stride_consts,
              const void* y0, int32_t stride_y0, void* out, void* stream);

/* ---- K5: surrogate (rate-matching) log-likelihood + gradient, no solve ----------------------
 * Replaces value_and_grad of the numpyro model in Modes.SURROGATE -- catalax/mcmc/mcmc.py:819-847
 * (sim_func = vmap(rate_fun) at the measured states, "{state}_init" = first state of each
 * measurement, data = surrogate rates), :469-480 (sigma_total[p,j] = sum_k J[p,j,k]^2 sigma_y[k]^2
 * + extra + 1e-8) and :486-491 (masked likelihood).  Points are grouped per measurement:
 * p = m * n_time + j.  t [N], y [N,S], theta [CH,P], consts [M,C], inits [M,S], data [N,S],
 * mask [N,S] or NULL (= isfinite(data)), jac2 [N,S,S] = J^2, extra2 [N,S] or NULL
 * (sigma_surr^2 + rate_sigma^2), sigma [CH,S] -> out [CH, 1+P+S] = (log L, d/dtheta, d/dsigma_y).
 * The model must have been generated with all parameter directions (as for ctx_loglik_grad). */
int ctx_rate_loglik_grad(ctx_model* m, int32_t likelihood, int64_t n_chains, int32_t n_points,
                         int32_t n_time, const void* t, const void* y, const void* theta,
                         const void* consts, const void* inits, const void* data,
                         const uint8_t* mask, const void* jac2, const void* extra2,
                         const void* sigma, void* out, void* stream);

/* ---- NeuralODE training: value and gradient w.r.t. the MLP weights ---------------------------
 * Replaces eqx.filter_value_and_grad of the training loss -- catalax/neural/trainer.py:268-306
 * (y_pred = vmap(model)(ti, y0i); loss(y_pred, yi)) through diffrax's discretise-then-optimise
 * adjoint of neuralode.py:38-59 / rateflow.py:100-121 (Tsit5, t0 = ts[0], SaveAt(ts)) and
 * universalode.py:118-141 (t0 = 0; grad_theta then also carries d/d gate_matrix,
 * d/d softplus(alpha_residual) and d/d mechanistic parameters at their offsets in theta).  Two calls around the caller's
 * loss: forward integrates, writes ys [B,T,S] and records every accepted step in the workspace;
 * backward takes gys = d loss / d ys [B,T,S] and the SAME workspace and returns
 * grad_theta [n_params] (flat parameter layout of the model) and optionally grad_y0 [B,S].
 * theta [n_params] is shared by the batch; ts is [B,T] (ts_per_row = 1) or [T].  `cap` = checkpoint
 * slots (accepted steps) per trajectory; a trajectory that needs more gets status 3.
 * dt0 <= 0 selects the reference's default, ts[1] - ts[0] of each trajectory. */
typedef struct ctx_adjoint_cfg {
  int64_t batch;
  int32_t n_times;
  int32_t max_steps;
  int32_t cap;
  int32_t ts_per_row;
  double rtol, atol, dt0;
  int32_t solver;      /* CTX_TSIT5 | CTX_DOPRI5 (trainer.py:128-132 lets the caller choose) */
  int32_t reserved;
} ctx_adjoint_cfg;
size_t ctx_mlp_adjoint_workspace_bytes(ctx_model* m, const ctx_adjoint_cfg* cfg);
int ctx_mlp_adjoint_forward(ctx_model* m, const ctx_adjoint_cfg* cfg, const void* y0,
                            const void* theta, const void* ts, void* ys, int32_t* status,
                            int32_t* n_accepted, int32_t* n_rejected, void* workspace,
                            size_t ws_bytes, void* stream);
int ctx_mlp_adjoint_backward(ctx_model* m, const ctx_adjoint_cfg* cfg, const void* y0,
                             const void* theta, const void* ts, const void* gys, int32_t* status,
                             void* grad_theta, void* grad_y0, void* workspace, size_t ws_bytes,
                             void* stream);

/* ---- host-buffer convenience (the call a reference-side ctypes binding would make) --------
 * Same as ctx_solve but every array is a HOST pointer; device staging, the H2D / D2H copies
 * and one stream synchronisation happen inside. */
int ctx_solve_host(ctx_model* m, const ctx_solve_cfg* cfg, const void* y0, const void* theta,
                   const void* consts, const void* ts, void* ys, int32_t* status,
                   int32_t* n_accepted, int32_t* n_rejected);

/* Same for ctx_loglik_grad: every array is a HOST pointer (partial / status may be NULL); one
 * NUTS leapfrog evaluation through this call moves theta [CH,P] + sigma [CH,S] (+ the constant
 * data block) to the device and out [CH, 1+D_theta+S] back.  Replaces the host-side round trip
 * of numpyro's potential_fn around mcmc.py:405-491. */
int ctx_loglik_grad_host(ctx_model* m, const ctx_loglik_cfg* cfg, const void* y0s,
                         const void* theta, const void* consts, const void* ts, const void* data,
                         const uint8_t* mask, const void* sigma, void* out, void* partial,
                         int32_t* status);

/* ---- XLA GPU custom-call shims (legacy "API_VERSION_ORIGINAL" ABI: no XLA/jaxlib header) -----
 * The reference calls its seam (simulation.py:160-196) from inside jit-compiled programs
 * (model.py:699, mcmc.py:448 under numpyro's jitted potential).  These entry points let XLA call
 * the kernels on ITS stream with ITS buffers: `buffers` = operands in order, then results in
 * order (the workspace is declared as the last result); `opaque` = one of the packed descriptors
 * below, byte-copied.  `stream` is a cudaStream_t.  Python registration: INTEGRATION.md section 3.
 *   ctx_xla_solve        operands y0, theta, consts, ts
 *                        results  ys, status, n_accepted, n_rejected, workspace
 *   ctx_xla_solve_sens   operands y0, theta, consts, ts
 *                        results  ys, sens, status, n_accepted, n_rejected, workspace
 *   ctx_xla_loglik_grad  operands y0s, theta, consts, ts, data, mask(uint8), sigma
 *                        results  out, partial, status, workspace
 * The ABI has no status channel: a failing call stores its code / message in a process-wide
 * slot; ctx_xla_last_status() returns and clears it (0 = no failure since the last query). */
typedef struct ctx_xla_solve_desc {
  int32_t abi;        /* CTX_ABI_VERSION */
  int32_t reserved;
  uint64_t model;     /* ctx_model* of this process */
  uint64_t ws_bytes;  /* size of the workspace result = ctx_workspace_bytes(model, &cfg) */
  ctx_solve_cfg cfg;
} ctx_xla_solve_desc;
typedef struct ctx_xla_loglik_desc {
  int32_t abi;
  int32_t has_mask;   /* 0: the mask operand is a dummy, use isfinite(data) (mcmc.py:566) */
  uint64_t model;
  uint64_t ws_bytes;  /* ctx_loglik_workspace_bytes(model, &cfg) */
  ctx_loglik_cfg cfg;
} ctx_xla_loglik_desc;
void ctx_xla_solve(void* stream, void** buffers, const char* opaque, size_t opaque_len);
void ctx_xla_solve_sens(void* stream, void** buffers, const char* opaque, size_t opaque_len);
void ctx_xla_loglik_grad(void* stream, void** buffers, const char* opaque, size_t opaque_len);
size_t ctx_xla_desc_bytes(int32_t which); /* 0: ctx_xla_solve_desc, 1: ctx_xla_loglik_desc */
int ctx_xla_last_status(void);
size_t ctx_xla_last_error(char* buf, size_t n);

/* ---- the path's only collective: one-shot all-reduce(sum) over NVLink peer memory --------------
 * Series-sharded log-density (every rank holds M/G series of ONE posterior): the [chains,
 * 1+D_theta+S] blocks of ctx_loglik_grad are summed over the ranks of the node once per evaluation
 * (mcmc.py:486-491 sums over the series inside one program).  ctx_p2p_create allocates this rank's
 * symmetric buffer and returns its 64-byte cudaIpc handle; the caller gathers all ranks' handles
 * (any out-of-band channel: torch.distributed, MPI) and passes them, in rank order, to
 * ctx_p2p_connect.  ctx_p2p_allreduce then sums `inout` (n elements of the model dtype, device
 * pointer) over the ranks in place with ONE kernel: copy into the symmetric buffer, flag exchange
 * (release / acquire at system scope), P2P loads of every peer's block summed in rank order (all
 * ranks get identical bits).  Stream-ordered, CUDA-graph capturable (the epoch lives in device
 * memory), at most 16 ranks of one node; every rank must call it the same number of times. */
typedef struct ctx_p2p ctx_p2p;
int ctx_p2p_create(int32_t rank, int32_t world, size_t bytes, ctx_p2p** out, void* handle_out);
int ctx_p2p_connect(ctx_p2p* p, const void* all_handles);
int ctx_p2p_allreduce(ctx_p2p* p, int32_t dtype, size_t n, void* inout, void* stream);
void ctx_p2p_free(ctx_p2p* p);
size_t ctx_p2p_last_error(char* buf, size_t n);

/* ---- diagnostics ------------------------------------------------------------------------ */
int ctx_abi_version(void);
int ctx_device_count(void); /* 0 when no GPU is visible; never throws */
size_t ctx_last_error(char* buf, size_t n);
/* Measured dense FMA peak of the current device in TFLOP/s (2 flops per FMA; FFMA for CTX_F32,
 * DFMA for CTX_F64): the roofline denominator of the compute-bound kernels. */
int ctx_measure_fma_peak(int32_t dtype, double* tflops);
/* number of kernel launches this library has issued in this process (bench "gpu_launches") */
int64_t ctx_launch_count(void);
/* name of the kernel the last ctx_solve / ctx_solve_sens call of THIS thread launched
 * ("ctx_solve_v0|v1|v2|w|mlp_tc"): lets tests assert which form `kernel = 0` selected */
const char* ctx_last_solve_kernel(void);
/* hash of the sources this library was built from (checked against the tree at load time) */
const char* ctx_build_stamp(void);

#ifdef __cplusplus
}
#endif
#endif /* CTX_B200_H_ */
