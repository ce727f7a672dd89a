// ctx_api.cu -- C ABI of libctxb200.so: NVRTC compilation of per-model kernels + launches.
//
// Boundary being replaced: the Python callable of catalax/tools/simulation.py:160-196
// (`Simulation._prepare_func`) -- see include/ctx_b200.h for the per-function citations.
// No torch types, no CPU fallback: without a CUDA device every compute entry point fails with
// CTX_E_NOGPU.
#include "../../include/ctx_b200.h"

#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <unistd.h>

extern const char ctx_integrator_src[];  // generated from ctx_integrator.cuh (embed step)
extern const char ctx_mlp_tc_src[];      // generated from ctx_mlp_tc.cuh
extern const char ctx_mlp_adjoint_src[]; // generated from ctx_mlp_adjoint.cuh
extern const char ctx_solve_v2_src[];    // generated from ctx_solve_v2.cuh
extern "C" int ctx_sched_heavy_first(const int* nacc, const int* nrej, int cap, long long B, int* order,
                                     int* scratch, void* stream);   // ctx_sched.cu

namespace {

thread_local std::string g_err;
thread_local const char* g_last_kernel = "";   // name of the solve kernel this thread launched last
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CTX_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver              \
                      ? CTX_E_NOGPU : CTX_E_CUDA,                                           \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                      \
  } while (0)

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}

std::string cache_dir() {
  const char* e = getenv("CTX_B200_CACHE");
  std::string d;
  if (e && *e) d = e;
  else {
    const char* home = getenv("HOME");
    d = std::string(home && *home ? home : "/tmp") + "/.cache/catalax_b200";
  }
  return d;
}

void mkdirs(const std::string& d) {
  std::string cur;
  for (size_t i = 0; i <= d.size(); ++i) {
    if (i == d.size() || d[i] == '/') {
      if (!cur.empty()) mkdir(cur.c_str(), 0755);
    }
    if (i < d.size()) cur += d[i];
  }
}

struct Variant {
  std::vector<char> cubin;
  cudaLibrary_t lib = nullptr;
  std::map<std::string, cudaKernel_t> kernels;
  int loaded_device = -1;
  long v1_warp_smem = -1;   // sizeof(ctx_warp_smem) read back from the loaded module
  long ll_warp_smem = -1;   // sizeof(ctx_ll_warp_smem) (cooperative log-likelihood kernel)
};

}  // namespace

struct ctx_model {
  std::string field_src;
  ctx_model_desc desc;
  std::mutex mu;
  std::map<int, Variant> variants;  // key = solver*2 + tan (+ 8 for the 8-point-group save variant)
  bool theta_ptr = false;           // generated with CTX_THETA_PTR (shared weights in smem)
  int p_pad = 0;
  int mlp_tc_smem = 0;              // > 0: the tcgen05 MLP kernel is available (bytes of smem)
  bool has_adjoint = false;         // weight-gradient kernels were generated (ctx_mlp_adjoint.cuh)
  int warp_warps = 0;               // > 0: lane-parallel form present (ctx_solve_w), warps per CTA
  // host-call staging (ctx_solve_host / ctx_loglik_grad_host): device buffers, two private
  // non-blocking streams and the events of the double-buffered pipeline.  `host_mu` is held for
  // the whole call, so concurrent host-buffer calls on ONE model serialise (device-pointer entry
  // points stay lock-free and stream-ordered).
  std::mutex host_mu;
  void* stage = nullptr;
  size_t stage_bytes = 0;
  int stage_device = -1;
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_solved[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
};

namespace {

// Kernel-argument block; layout must match `struct ctx_solve_args` in ctx_integrator.cuh.
template <typename real>
struct SolveArgs {
  const real* y0;
  const real* theta;
  const real* consts;
  const real* ts;
  real* ys;
  real* sens;
  int* status;
  int* n_accepted;
  int* n_rejected;
  unsigned long long* work_counter;
  long long B;
  int T;
  int max_steps;
  int stride_y0, stride_theta, stride_consts, stride_ts;
  int t0_from_ts;
  int inner;
  int q_div;
  int ts_smem;
  real rtol, atol, dt0;
  const real* ll_data;     // CTX_LLSAVE builds (ctx_pointwise_loglik)
  const real* ll_sigma;
  int ll_likelihood;
  const int* order;        // scheduling hint (cfg->order)
};

template <typename real>
struct RatesArgs {
  const real* t;
  const real* y;
  const real* theta;
  const real* consts;
  const real* y0;
  real* out;
  long long N;
  int stride_t, stride_theta, stride_consts, stride_y0;
};

template <typename real>
struct RateLlArgs {
  const real* t;
  const real* y;
  const real* theta;
  const real* consts;
  const real* inits;
  const real* data;
  const unsigned char* mask;
  const real* jac2;
  const real* extra2;
  const real* sigma;
  real* out;
  long long n_chains;
  int N, n_time, likelihood;
};

// Upper bound of sizeof(ctx_warp_smem) in ctx_integrator.cuh (input queue + step records +
// staging); used to choose CTX_WARPS before compiling.  The launch uses the exact size the
// module exports (ctx_v1_warp_smem_bytes).
size_t v1_warp_smem(const ctx_model_desc& d, int solver, int tan, bool theta_ptr = false) {
  const size_t w = d.dtype == CTX_F64 ? 8 : 4;
  const size_t S = (size_t)d.n_states;
  const size_t N = S * (1 + (tan ? d.n_dirs : 0));
  const size_t deg = solver <= 1 ? 4 : 1;
  const size_t qw = S + (theta_ptr ? 0 : (size_t)d.n_params) + (size_t)d.n_consts + 2;
  const size_t q_in = 2 * qw * 32 * w + 2 * 32 * 4;   // operands + trajectory indices
  if (!tan) {   // ctx_rec records (16-byte aligned) + one pending group per lane
    const size_t rec = (((2 + (deg + 1) * N) * w + 7 * 4) + 15) / 16 * 16;
    return q_in + 32 * rec + 32 * 8 * S * w + 64;   // (pending groups of up to 8 points)
  }
  const size_t rec = 2 + (deg + 1) * N;
  return q_in + 32 * rec * w + 3 * 32 * 8 + 32 * S * w + 32 * (N - S) * w + 64;
}

// Warps per CTA of the persistent kernel: 4 when that fits the default 48 KB, otherwise as many
// as fit ~96 KB (opt-in dynamic shared memory, two CTAs per SM); 0 = a single warp's records do
// not fit 200 KB, use the static kernel.
int v1_warps(const ctx_model_desc& d, int solver, int tan, size_t fixed = 0) {
  const size_t per_warp = v1_warp_smem(d, solver, tan, fixed > 0);
  if (fixed + 4 * per_warp <= 48 * 1024) return 4;
  if (fixed >= 96 * 1024) return per_warp + fixed <= 200 * 1024 ? 1 : 0;
  size_t warps = (96 * 1024 - fixed) / per_warp;
  if (warps > 4) warps = 4;
  if (warps == 0 && per_warp + fixed <= 200 * 1024) warps = 1;
  return (int)warps;
}

// NVRTC is loaded explicitly, by path, into a private namespace: the process usually holds
// another libnvrtc.so.12 already (PyTorch bundles the one of ITS CUDA release, 12.8 here), and a
// link-time dependency would bind to whichever copy was loaded first.  The kernels use PTX of
// the toolkit this library was built with (256-bit st.global.v8.f32 needs PTX ISA 8.8 = CUDA 12.9),
// so the toolkit's own NVRTC is preferred; with an older NVRTC those paths are compiled out.
struct NvrtcApi {
  void* handle = nullptr;
  nvrtcResult (*Version)(int*, int*) = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*,
                               const char* const*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  int major = 0, minor = 0;
  std::string path;
};

const NvrtcApi* nvrtc_api() {
  static NvrtcApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<std::string> cand;
    if (const char* e = getenv("CTX_B200_NVRTC")) cand.push_back(e);
    cand.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    cand.push_back("/usr/local/cuda/lib64/libnvrtc.so");
    cand.push_back("libnvrtc.so.12");
    cand.push_back("libnvrtc.so");
    for (const auto& c : cand) {
      void* h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
      if (!h) continue;
      NvrtcApi a;
      a.handle = h;
      a.path = c;
#define CTX_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name))
      CTX_SYM(Version, "nvrtcVersion");
      CTX_SYM(CreateProgram, "nvrtcCreateProgram");
      CTX_SYM(CompileProgram, "nvrtcCompileProgram");
      CTX_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
      CTX_SYM(GetProgramLog, "nvrtcGetProgramLog");
      CTX_SYM(DestroyProgram, "nvrtcDestroyProgram");
      CTX_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
      CTX_SYM(GetCUBIN, "nvrtcGetCUBIN");
      CTX_SYM(GetErrorString, "nvrtcGetErrorString");
#undef CTX_SYM
      if (!a.Version || !a.CreateProgram || !a.CompileProgram || !a.GetProgramLogSize ||
          !a.GetProgramLog || !a.DestroyProgram || !a.GetCUBINSize || !a.GetCUBIN ||
          !a.GetErrorString) {
        dlclose(h);
        continue;
      }
      a.Version(&a.major, &a.minor);
      api = a;
      break;
    }
  });
  return api.handle ? &api : nullptr;
}

size_t theta_smem(const ctx_model* m) {
  return m->theta_ptr ? (size_t)m->p_pad * (m->desc.dtype == CTX_F64 ? 8 : 4) : 0;
}

int compile_variant(ctx_model* m, int solver, int tan, Variant& v, int wide = 0, int ll = 0,
                    int direct = 0) {
  if (!v.cubin.empty()) return CTX_OK;
  std::string src;
  src += m->desc.dtype == CTX_F64 ? "typedef double real;\n#define R(x) x\n"
                                  : "typedef float real;\n#define R(x) x##f\n";
  src += "__device__ __forceinline__ real ctx_sign(real x) { return x > R(0.0) ? R(1.0) : "
         "(x < R(0.0) ? R(-1.0) : x); }\n";
  // divisions of the generated field (the code generator prints CTX_DIV(a, b)): IEEE, or for f32
  // builds with CTX_FAST_DIV bit 1 the SFU form (see ctx_integrator.cuh)
  src += m->desc.dtype == CTX_F64
             ? "#define CTX_DIV(a, b) ((a) / (b))\n"
             : "#if defined(CTX_FAST_DIV) && (CTX_FAST_DIV & 2)\n#define CTX_DIV(a, b) __fdividef((a), (b))\n"
               "#else\n#define CTX_DIV(a, b) ((a) / (b))\n#endif\n";
  src += m->field_src;
  src += "\n#include \"ctx_integrator.cuh\"\n";
  std::vector<std::string> opts = {
      "--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo",
      std::string("-DCTX_F64=") + (m->desc.dtype == CTX_F64 ? "1" : "0"),
      "-DCTX_SOLVER=" + std::to_string(solver), "-DCTX_TAN=" + std::to_string(tan),
      "-DCTX_LOG2S=" + std::to_string(std::log2((double)std::max(1, m->desc.n_states))) + "f",
      "-DCTX_WARPS=" + std::to_string(std::max(1, v1_warps(m->desc, solver, tan, theta_smem(m))))};
  if (v1_warps(m->desc, solver, tan, theta_smem(m)) <= 0) opts.push_back("-DCTX_NO_V1=1");
  {
    // Register cap of the persistent kernel (__launch_bounds__ min blocks).  Measured on the
    // 3-state model: 4 CTAs/SM (128 regs) f32 and 3 CTAs/SM (170 regs) f64 are the sweet spots;
    // tighter caps spill.  Budget ~96 registers + 10 per 32-bit word of trajectory state.
    const size_t words = (size_t)m->desc.n_states * (1 + (tan ? m->desc.n_dirs : 0)) *
                         (m->desc.dtype == CTX_F64 ? 2 : 1);
    const char* extra_probe = getenv("CTX_B200_NVRTC_FLAGS");
    size_t mb = 512 / (96 + 10 * words);
    if (wide) {
      // 8 requested times per lane and pass instead of 4 (CTX_GP=8): long output grids in f32
      // (measured, config 2: T=1000 4.19 -> 3.73 ms; T<=250 and f64 lose, see launch_solve).
      // 128 registers for the 3-state model: still 4 CTAs/SM.
      opts.push_back("-DCTX_GP=8");
    }
    // the dense-output pass writes point-wise log-likelihoods instead of states (ctx_pointwise_loglik)
    if (ll) opts.push_back("-DCTX_LLSAVE=1");
    // sparse output grids: the owning lane saves its own points (no cooperative pass)
    if (direct) opts.push_back("-DCTX_DIRECT_SAVE=1");
    mb = std::max<size_t>(1, std::min<size_t>(4, mb));
    if (m->theta_ptr) mb = 1;
    if (tan && !m->theta_ptr) {
      // log-likelihood kernels: a tighter cap than the solve kernel of the same build (measured on
      // config 4: f32 4 CTAs/SM +14 %, f64 3 CTAs/SM +4 %, 5 / 4 lose; see ctx_integrator.cuh)
      size_t llmb = std::max<size_t>(1, std::min<size_t>(4, 640 / (64 + 8 * words)));
      if (!(extra_probe && strstr(extra_probe, "-DCTX_LL_MIN_BLOCKS=")))
        opts.push_back("-DCTX_LL_MIN_BLOCKS=" + std::to_string(std::max(llmb, mb)));
    }
    const char* ex = getenv("CTX_B200_NVRTC_FLAGS");
    if (!(ex && strstr(ex, "-DCTX_MIN_BLOCKS=")))   // developer override for A/B runs
      opts.push_back("-DCTX_MIN_BLOCKS=" + std::to_string(mb));
  }
  const char* extra = getenv("CTX_B200_NVRTC_FLAGS");
  if (extra && *extra) {
    std::string e(extra), cur;
    for (char ch : e + " ") {
      if (ch == ' ') { if (!cur.empty()) opts.push_back(cur); cur.clear(); }
      else cur += ch;
    }
  }
  std::string keysrc = src + "\x01" + ctx_integrator_src + "\x01" + ctx_mlp_tc_src + "\x01" +
                       ctx_mlp_adjoint_src + "\x01" + ctx_solve_v2_src;
  for (auto& o : opts) keysrc += "\x02" + o;
  const NvrtcApi* nv = nvrtc_api();
  if (!nv) return fail(CTX_E_NVRTC, "libnvrtc.so.12 not found (set CTX_B200_NVRTC to its path)");
  const int nv_major = nv->major, nv_minor = nv->minor;
  if (nv_major < 12 || (nv_major == 12 && nv_minor < 9)) {
    opts.push_back("-DCTX_ST256=0");   // 256-bit vector stores need PTX ISA 8.8 (CUDA 12.9)
    keysrc += "\x02-DCTX_ST256=0";
  }
  keysrc += "\x03" + std::to_string(nv_major) + "." + std::to_string(nv_minor);
  char name[64];
  snprintf(name, sizeof name, "%016llx%016llx.cubin", (unsigned long long)fnv1a(keysrc),
           (unsigned long long)fnv1a(keysrc, 0x9e3779b97f4a7c15ull));
  const std::string dir = cache_dir();
  const std::string path = dir + "/" + name;
  const bool use_disk = !(getenv("CTX_B200_NO_DISK_CACHE"));
  if (use_disk) {
    if (FILE* f = fopen(path.c_str(), "rb")) {
      fseek(f, 0, SEEK_END);
      long n = ftell(f);
      fseek(f, 0, SEEK_SET);
      if (n > 0) {
        v.cubin.resize(n);
        if (fread(v.cubin.data(), 1, n, f) != (size_t)n) v.cubin.clear();
      }
      fclose(f);
      if (!v.cubin.empty()) return CTX_OK;
    }
  }
  nvrtcProgram prog;
  const char* hdr_src[] = {ctx_integrator_src, ctx_mlp_tc_src, ctx_mlp_adjoint_src, ctx_solve_v2_src};
  const char* hdr_name[] = {"ctx_integrator.cuh", "ctx_mlp_tc.cuh", "ctx_mlp_adjoint.cuh",
                            "ctx_solve_v2.cuh"};
  nvrtcResult r = nv->CreateProgram(&prog, src.c_str(), "ctx_model.cu", 4, hdr_src, hdr_name);
  if (r != NVRTC_SUCCESS) return fail(CTX_E_NVRTC, nv->GetErrorString(r));
  std::vector<const char*> copts;
  for (auto& o : opts) copts.push_back(o.c_str());
  r = nv->CompileProgram(prog, (int)copts.size(), copts.data());
  if (r != NVRTC_SUCCESS) {
    size_t ln = 0;
    nv->GetProgramLogSize(prog, &ln);
    std::string log(ln, '\0');
    nv->GetProgramLog(prog, log.data());
    nv->DestroyProgram(&prog);
    return fail(CTX_E_NVRTC, std::string("nvrtc (") + nv->path + "): " + nv->GetErrorString(r) + "\n" + log);
  }
  size_t n = 0;
  nv->GetCUBINSize(prog, &n);
  v.cubin.resize(n);
  nv->GetCUBIN(prog, v.cubin.data());
  nv->DestroyProgram(&prog);
  if (use_disk) {
    mkdirs(dir);
    std::string tmp = path + "." + std::to_string((long)getpid()) + ".tmp";
    if (FILE* f = fopen(tmp.c_str(), "wb")) {
      fwrite(v.cubin.data(), 1, v.cubin.size(), f);
      fclose(f);
      rename(tmp.c_str(), path.c_str());
    }
  }
  return CTX_OK;
}

int get_kernel(ctx_model* m, int solver, int tan, const char* kname, cudaKernel_t* out,
               int wide = 0, int ll = 0, int direct = 0) {
  if (solver < 0 || solver > 3) return fail(CTX_E_INVALID, "unknown solver id");
  if (tan && m->desc.n_dirs <= 0)
    return fail(CTX_E_INVALID, "model was generated without sensitivity directions");
  std::lock_guard<std::mutex> lk(m->mu);
  Variant& v = m->variants[solver * 2 + tan + (wide ? 8 : 0) + (ll ? 16 : 0) + (direct ? 32 : 0)];
  int rc = compile_variant(m, solver, tan, v, wide, ll, direct);
  if (rc) return rc;
  int dev = 0;
  CTX_CUDA(cudaGetDevice(&dev));
  if (!v.lib || v.loaded_device != dev) {
    CTX_CUDA(cudaLibraryLoadData(&v.lib, v.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    v.loaded_device = dev;
    v.kernels.clear();
  }
  auto it = v.kernels.find(kname);
  if (it == v.kernels.end()) {
    cudaKernel_t k;
    CTX_CUDA(cudaLibraryGetKernel(&k, v.lib, kname));
    it = v.kernels.emplace(kname, k).first;
  }
  *out = it->second;
  return CTX_OK;
}

// sizeof(ctx_warp_smem) of the loaded ctx_solve_v1 module (call after get_kernel)
int v1_exact_warp_smem(ctx_model* m, int solver, int tan, size_t* out, int wide = 0, int ll = 0,
                       int direct = 0) {
  std::lock_guard<std::mutex> lk(m->mu);
  Variant& v = m->variants[solver * 2 + tan + (wide ? 8 : 0) + (ll ? 16 : 0) + (direct ? 32 : 0)];
  if (v.v1_warp_smem < 0) {
    void* dptr = nullptr;
    size_t nbytes = 0;
    unsigned val = 0;
    CTX_CUDA(cudaLibraryGetGlobal(&dptr, &nbytes, v.lib, "ctx_v1_warp_smem_bytes"));
    CTX_CUDA(cudaMemcpy(&val, dptr, sizeof(val), cudaMemcpyDeviceToHost));
    v.v1_warp_smem = (long)val;
  }
  *out = (size_t)v.v1_warp_smem;
  return CTX_OK;
}

// sizeof(ctx_ll_warp_smem) of the loaded tangent module (call after get_kernel)
int ll_exact_warp_smem(ctx_model* m, int solver, size_t* out) {
  std::lock_guard<std::mutex> lk(m->mu);
  Variant& v = m->variants[solver * 2 + 1];
  if (v.ll_warp_smem < 0) {
    void* dptr = nullptr;
    size_t nbytes = 0;
    unsigned val = 0;
    CTX_CUDA(cudaLibraryGetGlobal(&dptr, &nbytes, v.lib, "ctx_ll_warp_smem_bytes"));
    CTX_CUDA(cudaMemcpy(&val, dptr, sizeof(val), cudaMemcpyDeviceToHost));
    v.ll_warp_smem = (long)val;
  }
  *out = (size_t)v.ll_warp_smem;
  return CTX_OK;
}

template <typename real>
int launch_solve(ctx_model* m, const ctx_solve_cfg* cfg, int tan, const void* y0,
                 const void* theta, const void* consts, const void* ts, void* ys, void* sens,
                 int32_t* status, int32_t* nacc, int32_t* nrej, void* ws, size_t ws_bytes,
                 cudaStream_t stream, const void* ll_data = nullptr, const void* ll_sigma = nullptr,
                 int ll_likelihood = 0, int probe_ts_stride = -1) {
  // probe_ts_stride >= 0: this IS the scheduling probe (one requested time per row, taken from the
  // caller's grid with this row stride); it must not start another one
  const ctx_model_desc& d = m->desc;
  if (cfg->batch <= 0) return CTX_OK;
  if (cfg->n_times <= 0) return fail(CTX_E_INVALID, "n_times must be >= 1");
  const size_t need = ctx_workspace_bytes(m, cfg);
  if (ws_bytes < need || (need && !ws)) return fail(CTX_E_WORKSPACE, "workspace too small");
  int* scratch = (int*)ws;  // [3*B] dummies when the caller passes NULL outputs
  SolveArgs<real> a;
  a.y0 = (const real*)y0;
  a.theta = (const real*)theta;
  a.consts = (const real*)consts;
  a.ts = (const real*)ts;
  a.ys = (real*)ys;
  a.sens = (real*)sens;
  unsigned long long* counter = (unsigned long long*)ws;
  char* p = (char*)ws + 256;
  a.status = status ? status : (int*)p;
  p += status ? 0 : sizeof(int) * cfg->batch;
  a.n_accepted = nacc ? nacc : (int*)p;
  p += nacc ? 0 : sizeof(int) * cfg->batch;
  a.n_rejected = nrej ? nrej : (int*)p;
  (void)scratch;
  a.work_counter = counter;
  a.B = cfg->batch;
  a.T = cfg->n_times;
  a.max_steps = cfg->max_steps;
  a.stride_y0 = cfg->in_axes[0] == 0 ? d.n_states : 0;
  a.stride_theta = cfg->in_axes[1] == 0 ? d.n_params : 0;
  a.stride_consts = cfg->in_axes[2] == 0 ? d.n_consts : 0;
  a.stride_ts = cfg->in_axes[3] == 0 ? cfg->n_times : 0;
  if (probe_ts_stride >= 0) a.stride_ts = probe_ts_stride;
  a.t0_from_ts = cfg->t0_from_ts ? 1 : 0;
  a.inner = cfg->inner > 0 ? cfg->inner : 0;
  a.q_div = 1;
  a.ts_smem = 0;
  if (a.inner > 0 && cfg->batch % a.inner != 0)
    return fail(CTX_E_INVALID, "two-level batch: batch must be a multiple of inner");
  a.rtol = (real)cfg->rtol;
  a.atol = (real)cfg->atol;
  a.dt0 = (real)cfg->dt0;
  a.ll_data = (const real*)ll_data;
  a.ll_sigma = (const real*)ll_sigma;
  a.ll_likelihood = ll_likelihood;
  a.order = (const int*)cfg->order;
  const int ll = ll_data ? 1 : 0;

  int kernel = cfg->kernel;
  bool auto_v2 = false;
  const int warps = v1_warps(m->desc, cfg->solver, tan, theta_smem(m));
  if (ll) {
    // the fused log-likelihood epilogue lives in the persistent kernel's cooperative save pass
    if (tan || m->theta_ptr || warps <= 0 || (kernel != 0 && kernel != 2))
      return fail(CTX_E_INVALID, "point-wise log-likelihood: symbolic models whose state fits the "
                                 "persistent kernel (kernel 0 or 2), no tangents");
    kernel = 2;
  }
  if (kernel == 0) {
    // The persistent kernel wins while a trajectory's state fits registers (measured: S=3 f32
    // 1.6x at T=16, 3.3x at T=1000); with S*(1+D) = 32 it spills and the static kernel is 3x
    // faster (profiles/).  Weight-array (MLP) models always take it.
    const size_t n_regs = (size_t)d.n_states * (1 + (tan ? d.n_dirs : 0)) * (d.dtype == CTX_F64 ? 2 : 1);
    kernel = (warps > 0 && (n_regs <= 24 || m->theta_ptr)) ? 2 : 1;
    // ... and a warp per trajectory (lane = state) beats both once the generated field has a
    // lane-parallel form (config 3, 32 states: 2.4x over the static kernel).
    if (kernel == 1 && m->warp_warps > 0 && !tan) kernel = 4;
    // Long shared f32 output grids: the warp-specialised form (producer warps step, consumer
    // warps evaluate and store; ctx_solve_v2.cuh) -- bit-identical; measured on config 2 (v1 / v2
    // ms): T=128 2.12 / 2.13, 250 2.64 / 2.66, 384 2.70 / 2.51, 512 3.01 / 2.68, 1000 3.74 / 3.51,
    // 2000 5.71 / 5.42; it loses when stepping dominates (T=16: 1.73 / 1.87) and in f64 (80
    // registers per thread), so it is selected only where the output dominates.
    // (f64: 8 + 8 warps in one CTA per SM, T=1000 6.1 -> 5.76 ms; selected for long grids only)
    if (kernel == 2 && !tan && !m->theta_ptr && d.n_states <= 4 && a.inner == 0 &&
        cfg->solver <= CTX_DOPRI5 && a.stride_ts == 0 &&
        cfg->n_times >= (d.dtype == CTX_F32 ? 320 : 768) &&
        cfg->n_times <= 8192 /* the grid is staged in shared memory */ &&
        cfg->batch >= 65536 && cfg->max_steps <= 65536 && !getenv("CTX_B200_NO_V2")) {
      kernel = 5;
      auto_v2 = true;
    }
  }
  // ---- scheduling probe (ctx_sched.cu), OPT-IN (cfg->reserved bit 1 or CTX_B200_PROBE=1): no order
  // hint from the caller -> classify the rows with a capped, looser-tolerance solve of the same
  // batch and hand the flagged ones out first.  The probe's outputs land in the caller's own output
  // arrays (all of them are overwritten by the real launch, stream-ordered).  Off by default: on
  // the benchmark workload the long rows are long only AT the tight tolerance (a 632-step row
  // takes 5 steps at rtol 1e-2), so a probe loose enough to be cheap does not find them and one
  // tight enough to find them (rtol 1e-4: recall 0.97) costs 40 % of the solve -- DESIGN.md section 4.
  if ((kernel == 2 || kernel == 5) && probe_ts_stride < 0 && !tan && !ll && !m->theta_ptr && warps > 0 &&
      a.order == nullptr && cfg->solver <= CTX_DOPRI5 && !cfg->t0_from_ts &&
      ((cfg->reserved & 2) || getenv("CTX_B200_PROBE")) &&
      cfg->batch >= (1ll << 16) && cfg->batch <= (1ll << 22) && cfg->rtol > 0 && cfg->dt0 > 0) {
    // A/B knobs (development): CTX_B200_PROBE_CAP / CTX_B200_PROBE_RTOL
    const char* e_cap = getenv("CTX_B200_PROBE_CAP");
    const char* e_tol = getenv("CTX_B200_PROBE_RTOL");
    const int cap = e_cap ? std::max(2, atoi(e_cap)) : 32;
    const double probe_rtol = e_tol ? atof(e_tol) : std::min(1e-2, 100.0 * cfg->rtol);
    const double loosen = probe_rtol / cfg->rtol;
    ctx_solve_cfg pc = *cfg;
    pc.n_times = 1;
    pc.max_steps = cap;
    pc.rtol = probe_rtol;
    pc.atol = cfg->atol * loosen;
    pc.kernel = 2;
    pc.order = nullptr;
    const real* ts_probe = (const real*)ts + (cfg->n_times - 1);
    int rc = launch_solve<real>(m, &pc, 0, y0, theta, consts, ts_probe, ys, nullptr, a.status,
                                a.n_accepted, a.n_rejected, ws, ws_bytes, stream, nullptr, nullptr, 0,
                                a.stride_ts);
    if (rc) return rc;
    char* q = (char*)ws + 256 + 3 * sizeof(int32_t) * (size_t)cfg->batch;
    q = (char*)(((uintptr_t)q + 255) / 256 * 256);
    int* order_ws = (int*)q;
    int* scratch_ws = order_ws + cfg->batch;
    cudaError_t e = (cudaError_t)ctx_sched_heavy_first(a.n_accepted, a.n_rejected, cap, cfg->batch, order_ws,
                                                       scratch_ws, stream);
    if (e != cudaSuccess) return fail(CTX_E_CUDA, std::string("ctx_sched_heavy_first: ") + cudaGetErrorString(e));
    a.order = order_ws;
    g_launches += 3;
  }
  cudaKernel_t k;
  void* params[] = {&a};
  bool launched = false;
  if (kernel == 5) {
    // warp-specialised persistent kernel (ctx_solve_v2.cuh): producers step, consumers save
    if (tan || m->theta_ptr || cfg->solver > CTX_DOPRI5)
      return fail(CTX_E_INVALID, "warp-specialised kernel: symbolic models, Tsit5 / Dopri5, no tangents");
    if (a.stride_ts != 0)
      return fail(CTX_E_INVALID, "warp-specialised kernel: the time grid must be shared (in_axes[3] = None)");
    if (cfg->batch >= (1ll << 31))
      return fail(CTX_E_INVALID, "persistent kernel: batch must be < 2^31 rows per call");
    int wide = (d.dtype == CTX_F32 && d.n_states <= 4 && cfg->n_times >= 768) ? 1 : 0;
    if (const char* ev = getenv("CTX_B200_WIDE_SAVE")) wide = atoi(ev) ? 1 : 0;
    int rc = get_kernel(m, cfg->solver, 0, "ctx_solve_v2", &k, wide);
    if (rc) return rc;
    unsigned info[3] = {0, 0, 0};   // sizeof(v2_prod_smem), producer warps, consumer warps per CTA
    {
      std::lock_guard<std::mutex> lk(m->mu);
      Variant& v = m->variants[cfg->solver * 2 + (wide ? 8 : 0)];
      void* dptr = nullptr;
      size_t nbytes = 0;
      CTX_CUDA(cudaLibraryGetGlobal(&dptr, &nbytes, v.lib, "ctx_v2_info"));
      CTX_CUDA(cudaMemcpy(info, dptr, sizeof(info), cudaMemcpyDeviceToHost));
    }
    const unsigned np = info[1], nc = info[2];
    const size_t ts_bytes = ((((size_t)cfg->n_times + 7) & ~(size_t)7) * sizeof(real) + 15) & ~(size_t)15;
    const size_t smem = ts_bytes + (size_t)np * info[0];
    if (smem > 227 * 1024) {
      if (!auto_v2)
        return fail(CTX_E_INVALID, "warp-specialised kernel: time grid + step-record ring exceed shared memory");
      kernel = 2;   // selected automatically but it does not fit: the persistent kernel takes over
    } else {
    int dev = 0, sms = 0, per_sm = 0;
    CTX_CUDA(cudaGetDevice(&dev));
    CTX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (smem > 48 * 1024)
      CTX_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    const unsigned block = (np + nc) * 32;
    CTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, (int)block, smem));
    if (per_sm < 1) per_sm = 1;
    unsigned long long grid = (unsigned long long)sms * per_sm;
    const unsigned long long want = (cfg->batch + 32ull * np - 1) / (32ull * np);
    if (grid > want) grid = want;
    a.q_div = (int)std::min<unsigned long long>(1u << 30, 4ull * grid * np);
    a.ts_smem = 1;
    CTX_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), stream));
    CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(block), params, smem, stream));
    g_launches++;
    g_last_kernel = "ctx_solve_v2";
    launched = true;
    }
  }
  if (launched) {
    // done
  } else if (kernel == 2) {
    if (warps <= 0)
      return fail(CTX_E_INVALID, "persistent kernel: step record does not fit shared memory");
    if (cfg->batch >= (1ll << 31))
      return fail(CTX_E_INVALID, "persistent kernel: batch must be < 2^31 rows per call");
    // Long output grids in f32 take the variant whose lanes evaluate 8 requested times per pass
    // (half the segment look-ups / record loads per point, two independent Horner groups, whole
    // 96-byte = 3-sector stores).  Measured on B200, config 2: T=1000 4.19 -> 3.73 ms, but
    // T=250 2.68 -> 3.03 ms (more partial groups per segment) and f64 T=1000 6.1 -> 7.6 ms.
    int wide = (!tan && d.dtype == CTX_F32 && !m->theta_ptr && d.n_states <= 4 &&
                cfg->n_times >= 768) ? 1 : 0;
    if (const char* ev = getenv("CTX_B200_WIDE_SAVE")) wide = atoi(ev) ? (tan ? 0 : 1) : 0;
    // Sparse grids (few requested times per trajectory): the lane that owns a step saves its own
    // points; the cooperative pass only pays when steps cover many points (measured thresholds:
    // DESIGN.md section 4).  CTX_B200_DIRECT_SAVE=0|1 overrides.
    int direct = (!tan && !ll && !m->theta_ptr && cfg->n_times <= 32) ? 1 : 0;
    if (const char* ev = getenv("CTX_B200_DIRECT_SAVE")) direct = (atoi(ev) && !tan && !ll) ? 1 : 0;
    int rc = get_kernel(m, cfg->solver, tan, "ctx_solve_v1", &k, wide, ll, direct);
    if (rc) return rc;
    const unsigned block = 32u * warps;
    int dev = 0, sms = 0, per_sm = 0;
    CTX_CUDA(cudaGetDevice(&dev));
    CTX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    size_t per_warp = 0;
    rc = v1_exact_warp_smem(m, cfg->solver, tan, &per_warp, wide, ll, direct);
    if (rc) return rc;
    size_t smem = theta_smem(m) + (size_t)warps * per_warp;
    if (m->theta_ptr && a.stride_theta != 0)
      return fail(CTX_E_INVALID, "a weight-array model takes shared parameters (in_axes[1] = None)");
    // a time grid shared by all trajectories is staged in shared memory (<= 32 KB)
    const size_t ts_bytes = (((size_t)cfg->n_times + 7) & ~(size_t)7) * sizeof(real);
    if (a.stride_ts == 0 && ts_bytes <= 32 * 1024 && smem + ts_bytes <= 200 * 1024 &&
        !getenv("CTX_B200_NO_TS_SMEM")) {
      a.ts_smem = 1;
      smem += ts_bytes;
    }
    if (smem > 48 * 1024)
      CTX_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    CTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, (int)block,
                                                           smem));
    if (per_sm < 1) per_sm = 1;
    unsigned long long grid = (unsigned long long)sms * per_sm;
    const unsigned long long want = (cfg->batch + block - 1) / block;
    if (grid > want) grid = want;
    // guided self-scheduling of the input queue: a warp claims remaining / (4 * warps in the
    // grid) trajectories at a time, clamped to [1, 32]
    a.q_div = (int)std::min<unsigned long long>(1u << 30, 4ull * grid * (unsigned long long)warps);
    CTX_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), stream));
    CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(block), params, smem,
                              stream));
    g_launches++;
    g_last_kernel = ll ? "ctx_solve_v1 (log-likelihood save)" : (direct ? "ctx_solve_v1 (direct save)" : "ctx_solve_v1");
  } else if (kernel == 3) {
    // K3: tcgen05 MLP tiles (ctx_mlp_tc.cuh); theta = [aux | UMMA weight pack]
    if (m->mlp_tc_smem <= 0 || m->desc.dtype != CTX_F32 || tan || cfg->solver > CTX_DOPRI5)
      return fail(CTX_E_INVALID, "tensor-core MLP kernel: f32 MLP model with Tsit5/Dopri5 only");
    if (a.stride_theta != 0)
      return fail(CTX_E_INVALID, "a weight-array model takes shared parameters (in_axes[1] = None)");
    int rc = get_kernel(m, cfg->solver, 0, "ctx_solve_mlp_tc", &k);
    if (rc) return rc;
    int dev = 0, sms = 0;
    CTX_CUDA(cudaGetDevice(&dev));
    CTX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CTX_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  m->mlp_tc_smem));
    unsigned long long grid = (cfg->batch + 255) / 256;   // 2 tiles x 128 trajectories per CTA
    if (grid > (unsigned long long)sms) grid = sms;
    CTX_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), stream));
    CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(512), params,
                              (size_t)m->mlp_tc_smem, stream));
    g_launches++;
    g_last_kernel = "ctx_solve_mlp_tc";
  } else if (kernel == 4) {
    // K1w: one warp per trajectory, persistent warps (ctx_integrator.cuh ctx_solve_w)
    if (m->warp_warps <= 0 || tan)
      return fail(CTX_E_INVALID, "warp-per-trajectory kernel: model has no lane-parallel form "
                                 "(or tangents were requested)");
    int rc = get_kernel(m, cfg->solver, 0, "ctx_solve_w", &k);
    if (rc) return rc;
    const unsigned block = 32u * m->warp_warps;
    int dev = 0, sms = 0, per_sm = 0;
    CTX_CUDA(cudaGetDevice(&dev));
    CTX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, (int)block, 0));
    if (per_sm < 1) per_sm = 1;
    unsigned long long grid = (unsigned long long)sms * per_sm;
    const unsigned long long want = (cfg->batch + m->warp_warps - 1) / m->warp_warps;
    if (grid > want) grid = want;
    CTX_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), stream));
    CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(block), params, 0, stream));
    g_launches++;
    g_last_kernel = "ctx_solve_w";
  } else if (kernel == 1) {
    if (m->theta_ptr && warps > 0)
      return fail(CTX_E_INVALID, "weight-array models run the persistent kernel (kernel 0 or 2)");
    int rc = get_kernel(m, cfg->solver, tan, "ctx_solve_v0", &k);
    if (rc) return rc;
    unsigned block = 128;
    if (const char* e = getenv("CTX_B200_V0_BLOCK")) block = (unsigned)std::max(32, std::min(128, atoi(e)));
    const unsigned long long grid = (cfg->batch + block - 1) / block;
    CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(block), params, 0, stream));
    g_launches++;
    g_last_kernel = "ctx_solve_v0";
  } else {
    return fail(CTX_E_INVALID, "unknown kernel id");
  }
  return CTX_OK;
}

template <typename real>
struct LoglikArgs {
  const real* y0s;
  const real* theta;
  const real* consts;
  const real* ts;
  const real* data;
  const unsigned char* mask;
  const real* sigma;
  real* partial;
  int* status;
  int* n_accepted;
  int* n_rejected;
  unsigned long long* work_counter;
  long long n_traj;
  int M, T, likelihood, max_steps;
  real rtol, atol, dt0;
  const int* order;
};

template <typename real>
int launch_loglik(ctx_model* m, const ctx_loglik_cfg* cfg, const void* y0s, const void* theta,
                  const void* consts, const void* ts, const void* data, const uint8_t* mask,
                  const void* sigma, void* out, void* partial, int32_t* status, void* ws,
                  size_t ws_bytes, cudaStream_t stream) {
  const ctx_model_desc& d = m->desc;
  const long long n_traj = cfg->n_chains * (long long)cfg->n_series;
  if (n_traj <= 0) return CTX_OK;
  const size_t need = ctx_loglik_workspace_bytes(m, cfg);
  if (ws_bytes < need || !ws) return fail(CTX_E_WORKSPACE, "workspace too small");
  const int kp = 1 + d.n_theta_dirs + d.n_states + (d.n_dirs - d.n_theta_dirs);
  char* p = (char*)ws;
  LoglikArgs<real> a;
  a.work_counter = (unsigned long long*)p; p += 256;
  a.status = status ? status : (int*)p; p += sizeof(int) * n_traj;
  a.n_accepted = (int*)p; p += sizeof(int) * n_traj;
  a.n_rejected = (int*)p; p += sizeof(int) * n_traj;
  p = (char*)(((uintptr_t)p + 255) / 256 * 256);
  a.partial = partial ? (real*)partial : (real*)p;
  a.y0s = (const real*)y0s; a.theta = (const real*)theta; a.consts = (const real*)consts;
  a.ts = (const real*)ts; a.data = (const real*)data; a.mask = mask; a.sigma = (const real*)sigma;
  a.n_traj = n_traj; a.M = cfg->n_series; a.T = cfg->n_times; a.likelihood = cfg->likelihood;
  a.max_steps = cfg->max_steps;
  a.rtol = (real)cfg->rtol; a.atol = (real)cfg->atol; a.dt0 = (real)cfg->dt0;
  a.order = (const int*)cfg->order;
  // kernel (cfg->kernel): 0 = automatic, 1 = per-lane likelihood terms (ctx_loglik_v1),
  // 2 = warp-cooperative likelihood terms (ctx_loglik_v2, the default)
  const int which = cfg->kernel == 1 ? 1 : 2;
  if (cfg->kernel < 0 || cfg->kernel > 2) return fail(CTX_E_INVALID, "unknown log-likelihood kernel id");
  cudaKernel_t k;
  int rc = get_kernel(m, cfg->solver, 1, which == 1 ? "ctx_loglik_v1" : "ctx_loglik_v2", &k);
  if (rc) return rc;
  const int warps = std::max(1, v1_warps(d, cfg->solver, 1, theta_smem(m)));
  const unsigned block = 32u * warps;
  size_t smem = 0;
  if (which == 2) {
    if ((long long)cfg->n_series * cfg->n_times >= (1ll << 31))
      return fail(CTX_E_INVALID, "log-likelihood: n_series * n_times must be < 2^31");
    size_t per_warp = 0;
    rc = ll_exact_warp_smem(m, cfg->solver, &per_warp);
    if (rc) return rc;
    smem = (size_t)warps * per_warp;
    if (smem > 200 * 1024)
      return fail(CTX_E_INVALID, "log-likelihood: step records do not fit shared memory");
    if (smem > 48 * 1024)
      CTX_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  }
  int dev = 0, sms = 0, per_sm = 0;
  CTX_CUDA(cudaGetDevice(&dev));
  CTX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, (int)block, smem));
  if (per_sm < 1) per_sm = 1;
  unsigned long long grid = (unsigned long long)sms * per_sm;
  const unsigned long long want = (n_traj + block - 1) / block;
  if (grid > want) grid = want;
  void* params[] = {&a};
  CTX_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), stream));
  CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(block), params, smem, stream));
  g_launches++;
  cudaKernel_t kr;
  rc = get_kernel(m, cfg->solver, 1, "ctx_loglik_reduce", &kr);
  if (rc) return rc;
  const real* part = a.partial;
  real* outp = (real*)out;
  long long nch = cfg->n_chains;
  int M = cfg->n_series, KR = 1 + d.n_theta_dirs + d.n_states;
  void* rparams[] = {(void*)&part, (void*)&outp, (void*)&nch, (void*)&M, (void*)&KR};
  const unsigned long long rgrid = (nch * 32 + 127) / 128;
  CTX_CUDA(cudaLaunchKernel((const void*)kr, dim3((unsigned)rgrid), dim3(128), rparams, 0, stream));
  g_launches++;
  (void)kp;
  return CTX_OK;
}

// ---------------------------------------------------------------------------- MLP adjoint
template <typename real>
struct AdjArgs {
  const real* y0;
  const real* theta;
  const real* ts;
  real* ys;
  const real* gys;
  real* ck_real;
  int* ck_int;
  int* n_steps;
  int* status;
  int* n_accepted;
  int* n_rejected;
  real* partial;
  real* gy0;
  unsigned long long* work_counter;
  long long B;
  int T, cap, max_steps, stride_ts;
  real rtol, atol, dt0;
};

struct AdjPlan {
  unsigned spack = 0, nact = 0, warps = 0, n_params = 0, wmax = 0;   // ctx_adj_info of the module
  unsigned grid = 0;
  size_t smem_fwd = 0, smem_bwd = 0;
  size_t off_ck_real = 0, off_ck_int = 0, off_nsteps = 0, off_counts = 0, off_partial = 0, total = 0;
};

int adj_plan(ctx_model* m, const ctx_adjoint_cfg* cfg, AdjPlan& p) {
  if (!m->has_adjoint)
    return fail(CTX_E_INVALID, "model has no weight-gradient kernels (plain NeuralODE fields only)");
  if (cfg->batch <= 0 || cfg->n_times <= 0 || cfg->cap <= 0)
    return fail(CTX_E_INVALID, "adjoint: batch, n_times and cap must be positive");
  if (cfg->solver != CTX_TSIT5 && cfg->solver != CTX_DOPRI5)
    return fail(CTX_E_INVALID, "adjoint: Tsit5 or Dopri5");
  cudaKernel_t k;
  int rc = get_kernel(m, cfg->solver, 0, "ctx_mlp_adj_forward", &k);
  if (rc) return rc;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    Variant& v = m->variants[cfg->solver * 2 + 0];
    void* dptr = nullptr;
    size_t nbytes = 0;
    unsigned info[5] = {0, 0, 0, 0, 0};
    CTX_CUDA(cudaLibraryGetGlobal(&dptr, &nbytes, v.lib, "ctx_adj_info"));
    CTX_CUDA(cudaMemcpy(info, dptr, sizeof(info), cudaMemcpyDeviceToHost));
    p.spack = info[0]; p.nact = info[1]; p.warps = info[2]; p.n_params = info[3];
    p.wmax = info[4];
  }
  const size_t w = m->desc.dtype == CTX_F64 ? 8 : 4;
  int dev = 0, sms = 0;
  CTX_CUDA(cudaGetDevice(&dev));
  CTX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  unsigned long long grid = ((unsigned long long)cfg->batch + p.warps - 1) / p.warps;
  if (grid > (unsigned long long)sms) grid = sms;
  p.grid = (unsigned)grid;
  p.smem_fwd = ((size_t)p.spack + (size_t)p.warps * p.nact) * w;
  p.smem_bwd = ((size_t)p.spack + (size_t)p.warps * (p.spack + 7 * (size_t)p.nact + 2 * (size_t)p.wmax)) * w;
  auto al = [](size_t n) { return (n + 255) / 256 * 256; };
  const size_t B = (size_t)cfg->batch, cap = (size_t)cfg->cap, S = (size_t)m->desc.n_states;
  size_t off = 256;
  p.off_ck_real = off; off += al(B * cap * (2 + S) * w);
  p.off_ck_int = off;  off += al(B * cap * 2 * sizeof(int));
  p.off_nsteps = off;  off += al(B * sizeof(int));
  p.off_counts = off;  off += al(3 * B * sizeof(int));
  p.off_partial = off; off += al((size_t)p.grid * p.warps * p.n_params * w);
  p.total = off;
  return CTX_OK;
}

template <typename real>
int launch_adjoint(ctx_model* m, const ctx_adjoint_cfg* cfg, bool backward, const void* y0,
                   const void* theta, const void* ts, void* ys, const void* gys, int32_t* status,
                   int32_t* nacc, int32_t* nrej, void* grad_theta, void* grad_y0, void* ws,
                   size_t ws_bytes, cudaStream_t stream) {
  AdjPlan p;
  int rc = adj_plan(m, cfg, p);
  if (rc) return rc;
  if (!ws || ws_bytes < p.total) return fail(CTX_E_WORKSPACE, "workspace too small");
  char* base = (char*)ws;
  AdjArgs<real> a;
  a.y0 = (const real*)y0; a.theta = (const real*)theta; a.ts = (const real*)ts;
  a.ys = (real*)ys; a.gys = (const real*)gys;
  a.ck_real = (real*)(base + p.off_ck_real);
  a.ck_int = (int*)(base + p.off_ck_int);
  a.n_steps = (int*)(base + p.off_nsteps);
  int* counts = (int*)(base + p.off_counts);
  a.status = status ? status : counts;
  a.n_accepted = nacc ? nacc : counts + cfg->batch;
  a.n_rejected = nrej ? nrej : counts + 2 * cfg->batch;
  a.partial = (real*)(base + p.off_partial);
  a.gy0 = (real*)grad_y0;
  a.work_counter = (unsigned long long*)base;
  a.B = cfg->batch; a.T = cfg->n_times; a.cap = cfg->cap; a.max_steps = cfg->max_steps;
  a.stride_ts = cfg->ts_per_row ? cfg->n_times : 0;
  a.rtol = (real)cfg->rtol; a.atol = (real)cfg->atol; a.dt0 = (real)cfg->dt0;
  cudaKernel_t k;
  rc = get_kernel(m, cfg->solver, 0, backward ? "ctx_mlp_adj_backward" : "ctx_mlp_adj_forward", &k);
  if (rc) return rc;
  const size_t smem = backward ? p.smem_bwd : p.smem_fwd;
  if (smem > 227 * 1024) return fail(CTX_E_INVALID, "adjoint: the network does not fit shared memory");
  if (smem > 48 * 1024)
    CTX_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  void* params[] = {&a};
  CTX_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), stream));
  CTX_CUDA(cudaLaunchKernel((const void*)k, dim3(p.grid), dim3(32 * p.warps), params, smem, stream));
  g_launches++;
  if (backward) {
    cudaKernel_t kr;
    rc = get_kernel(m, cfg->solver, 0, "ctx_mlp_adj_reduce", &kr);
    if (rc) return rc;
    const real* part = a.partial;
    real* g = (real*)grad_theta;
    int rows = (int)(p.grid * p.warps);
    void* rparams[] = {(void*)&part, (void*)&g, (void*)&rows};
    CTX_CUDA(cudaLaunchKernel((const void*)kr, dim3((p.n_params + 127) / 128), dim3(128), rparams,
                              0, stream));
    g_launches++;
  }
  return CTX_OK;
}

int dispatch_solve(ctx_model* m, const ctx_solve_cfg* cfg, int tan, const void* y0,
                   const void* theta, const void* consts, const void* ts, void* ys, void* sens,
                   int32_t* status, int32_t* nacc, int32_t* nrej, void* ws, size_t ws_bytes,
                   void* stream) {
  if (!m || !cfg) return fail(CTX_E_INVALID, "null model or config");
  if (ctx_device_count() <= 0)
    return fail(CTX_E_NOGPU, "no CUDA device visible: catalax_b200 has no CPU fallback");
  if (m->desc.dtype == CTX_F64)
    return launch_solve<double>(m, cfg, tan, y0, theta, consts, ts, ys, sens, status, nacc, nrej,
                                ws, ws_bytes, (cudaStream_t)stream);
  return launch_solve<float>(m, cfg, tan, y0, theta, consts, ts, ys, sens, status, nacc, nrej, ws,
                             ws_bytes, (cudaStream_t)stream);
}

}  // namespace

extern "C" {

int ctx_abi_version(void) { return CTX_ABI_VERSION; }

int ctx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

size_t ctx_last_error(char* buf, size_t n) {
  if (buf && n) {
    size_t k = g_err.size() < n - 1 ? g_err.size() : n - 1;
    memcpy(buf, g_err.data(), k);
    buf[k] = '\0';
  }
  return g_err.size();
}

int64_t ctx_launch_count(void) { return g_launches.load(); }

const char* ctx_last_solve_kernel(void) { return g_last_kernel; }

int ctx_model_compile(const char* field_src, const ctx_model_desc* desc, ctx_model** out) {
  if (!field_src || !desc || !out) return fail(CTX_E_INVALID, "null argument");
  if (desc->n_states <= 0 || desc->n_params < 0 || desc->n_consts < 0 || desc->n_dirs < 0 ||
      (desc->dtype != CTX_F32 && desc->dtype != CTX_F64))
    return fail(CTX_E_INVALID, "bad model descriptor");
  // the generated text carries its own sizes; refuse a mismatching descriptor
  auto has = [&](const char* macro, int v) {
    std::string needle = std::string("#define ") + macro + " " + std::to_string(v) + "\n";
    return strstr(field_src, needle.c_str()) != nullptr;
  };
  if (!has("CTX_S", desc->n_states) || !has("CTX_P", desc->n_params) ||
      !has("CTX_C", desc->n_consts) || !has("CTX_D", desc->n_dirs))
    return fail(CTX_E_INVALID, "descriptor does not match the CTX_S/P/C/D macros of field_src");
  ctx_model* m = new ctx_model;
  m->field_src = field_src;
  m->desc = *desc;
  m->theta_ptr = strstr(field_src, "#define CTX_THETA_PTR 1") != nullptr;
  if (m->theta_ptr) {
    const char* q = strstr(field_src, "#define CTX_P_PAD ");
    m->p_pad = q ? atoi(q + 18) : desc->n_params;
    if (m->p_pad < desc->n_params) m->p_pad = desc->n_params;
  }
  if (const char* q = strstr(field_src, "#define CTX_MLP_TC_SMEM ")) m->mlp_tc_smem = atoi(q + 24);
  m->has_adjoint = strstr(field_src, "#define CTX_MLP_ADJ 1") != nullptr;
  if (strstr(field_src, "#define CTX_HAVE_WARP 1"))
    if (const char* q = strstr(field_src, "#define CTXW_WARPS ")) m->warp_warps = atoi(q + 19);
  *out = m;
  return CTX_OK;
}

void ctx_model_free(ctx_model* m) {
  if (!m) return;
  for (auto& kv : m->variants)
    if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
  if (m->stage) cudaFree(m->stage);
  for (int i = 0; i < 2; ++i) {
    if (m->ev_solved[i]) cudaEventDestroy(m->ev_solved[i]);
    if (m->ev_copied[i]) cudaEventDestroy(m->ev_copied[i]);
  }
  if (m->s_compute) cudaStreamDestroy(m->s_compute);
  if (m->s_copy) cudaStreamDestroy(m->s_copy);
  delete m;
}

int ctx_model_cubin(ctx_model* m, int32_t solver, int32_t with_tangents, const void** cubin,
                    size_t* nbytes) {
  if (!m || !cubin || !nbytes) return fail(CTX_E_INVALID, "null argument");
  if (solver < 0 || solver > 3) return fail(CTX_E_INVALID, "unknown solver id");
  if (with_tangents && m->desc.n_dirs <= 0)
    return fail(CTX_E_INVALID, "model was generated without sensitivity directions");
  std::lock_guard<std::mutex> lk(m->mu);
  Variant& v = m->variants[solver * 2 + (with_tangents ? 1 : 0)];
  int rc = compile_variant(m, solver, with_tangents ? 1 : 0, v);
  if (rc) return rc;
  *cubin = v.cubin.data();
  *nbytes = v.cubin.size();
  return CTX_OK;
}

size_t ctx_workspace_bytes(const ctx_model* m, const ctx_solve_cfg* cfg) {
  (void)m;
  if (!cfg || cfg->batch <= 0) return 0;
  // [256 B: work counter][3 x int32[B]: dummies for NULL status / counts][probe: order int32[B] +
  // (B / 1024 + 2) ints of scan scratch, 256-byte aligned]
  const size_t B = (size_t)cfg->batch;
  return 256 + 3 * sizeof(int32_t) * B + 256 + sizeof(int32_t) * B + sizeof(int32_t) * (B / 1024 + 2) + 256;
}

int ctx_solve(ctx_model* m, const ctx_solve_cfg* cfg, const void* y0, const void* theta,
              const void* consts, const void* ts, void* ys, int32_t* status, int32_t* nacc,
              int32_t* nrej, void* ws, size_t ws_bytes, void* stream) {
  return dispatch_solve(m, cfg, 0, y0, theta, consts, ts, ys, nullptr, status, nacc, nrej, ws,
                        ws_bytes, stream);
}

int ctx_solve_sens(ctx_model* m, const ctx_solve_cfg* cfg, const void* y0, const void* theta,
                   const void* consts, const void* ts, void* ys, void* sens, int32_t* status,
                   int32_t* nacc, int32_t* nrej, void* ws, size_t ws_bytes, void* stream) {
  if (!sens) return fail(CTX_E_INVALID, "sens output is null");
  return dispatch_solve(m, cfg, 1, y0, theta, consts, ts, ys, sens, status, nacc, nrej, ws,
                        ws_bytes, stream);
}

int ctx_pointwise_loglik(ctx_model* m, const ctx_solve_cfg* cfg, int32_t likelihood, const void* y0,
                         const void* theta, const void* consts, const void* ts, const void* data,
                         const void* sigma, void* ll, int32_t* status, int32_t* nacc, int32_t* nrej,
                         void* ws, size_t ws_bytes, void* stream) {
  if (!m || !cfg || !data || !sigma || !ll) return fail(CTX_E_INVALID, "null argument");
  if (likelihood != CTX_LIK_NORMAL && likelihood != CTX_LIK_SOFTLAPLACE)
    return fail(CTX_E_INVALID, "unknown likelihood id");
  if (ctx_device_count() <= 0)
    return fail(CTX_E_NOGPU, "no CUDA device visible: catalax_b200 has no CPU fallback");
  if (m->desc.dtype == CTX_F64)
    return launch_solve<double>(m, cfg, 0, y0, theta, consts, ts, ll, nullptr, status, nacc, nrej, ws,
                                ws_bytes, (cudaStream_t)stream, data, sigma, likelihood);
  return launch_solve<float>(m, cfg, 0, y0, theta, consts, ts, ll, nullptr, status, nacc, nrej, ws,
                             ws_bytes, (cudaStream_t)stream, data, sigma, likelihood);
}

size_t ctx_loglik_workspace_bytes(const ctx_model* m, const ctx_loglik_cfg* cfg) {
  if (!m || !cfg || cfg->n_chains <= 0 || cfg->n_series <= 0) return 0;
  const ctx_model_desc& d = m->desc;
  const size_t w = d.dtype == CTX_F64 ? 8 : 4;
  const size_t n_traj = (size_t)cfg->n_chains * cfg->n_series;
  const size_t kp = 1 + d.n_theta_dirs + d.n_states + (d.n_dirs - d.n_theta_dirs);
  return 256 + 3 * sizeof(int32_t) * n_traj + 256 + n_traj * kp * w;
}

int ctx_loglik_grad(ctx_model* m, const ctx_loglik_cfg* cfg, const void* y0s, const void* theta,
                    const void* consts, const void* ts, const void* data, const uint8_t* mask,
                    const void* sigma, void* out, void* partial, int32_t* status, void* ws,
                    size_t ws_bytes, void* stream) {
  if (!m || !cfg || !out) return fail(CTX_E_INVALID, "null argument");
  if (m->desc.n_theta_dirs != m->desc.n_params)
    return fail(CTX_E_INVALID, "model must be generated with all parameter directions");
  if (cfg->likelihood != CTX_LIK_NORMAL && cfg->likelihood != CTX_LIK_SOFTLAPLACE)
    return fail(CTX_E_INVALID, "unknown likelihood id");
  if (cfg->n_times <= 0) return fail(CTX_E_INVALID, "n_times must be >= 1");
  if (ctx_device_count() <= 0)
    return fail(CTX_E_NOGPU, "no CUDA device visible: catalax_b200 has no CPU fallback");
  if (m->desc.dtype == CTX_F64)
    return launch_loglik<double>(m, cfg, y0s, theta, consts, ts, data, mask, sigma, out, partial,
                                 status, ws, ws_bytes, (cudaStream_t)stream);
  return launch_loglik<float>(m, cfg, y0s, theta, consts, ts, data, mask, sigma, out, partial,
                              status, ws, ws_bytes, (cudaStream_t)stream);
}

int ctx_rates(ctx_model* m, int64_t n, const void* t, int32_t stride_t, const void* y,
              const void* theta, int32_t stride_theta, const void* consts, int32_t stride_consts,
              const void* y0, int32_t stride_y0, void* out, void* stream) {
  if (!m || !out) return fail(CTX_E_INVALID, "null argument");
  if (n <= 0) return CTX_OK;
  if (ctx_device_count() <= 0)
    return fail(CTX_E_NOGPU, "no CUDA device visible: catalax_b200 has no CPU fallback");
  cudaKernel_t k;
  int rc = get_kernel(m, CTX_TSIT5, 0, "ctx_rates", &k);
  if (rc) return rc;
  const unsigned block = 128;
  const unsigned long long grid = (n + block - 1) / block;
  if (m->desc.dtype == CTX_F64) {
    RatesArgs<double> a{(const double*)t, (const double*)y, (const double*)theta,
                        (const double*)consts, (const double*)y0, (double*)out, n,
                        stride_t, stride_theta, stride_consts, stride_y0};
    void* params[] = {&a};
    CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(block), params, 0,
                              (cudaStream_t)stream));
  } else {
    RatesArgs<float> a{(const float*)t, (const float*)y, (const float*)theta,
                       (const float*)consts, (const float*)y0, (float*)out, n,
                       stride_t, stride_theta, stride_consts, stride_y0};
    void* params[] = {&a};
    CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(block), params, 0,
                              (cudaStream_t)stream));
  }
  g_launches++;
  return CTX_OK;
}

int ctx_rate_loglik_grad(ctx_model* m, int32_t likelihood, int64_t n_chains, int32_t n_points,
                         int32_t n_time, const void* t, const void* y, const void* theta,
                         const void* consts, const void* inits, const void* data,
                         const uint8_t* mask, const void* jac2, const void* extra2,
                         const void* sigma, void* out, void* stream) {
  if (!m || !out) return fail(CTX_E_INVALID, "null argument");
  if (m->desc.n_theta_dirs != m->desc.n_params || m->desc.n_dirs <= 0)
    return fail(CTX_E_INVALID, "model must be generated with all parameter directions");
  if (likelihood != CTX_LIK_NORMAL && likelihood != CTX_LIK_SOFTLAPLACE)
    return fail(CTX_E_INVALID, "unknown likelihood id");
  if (n_time <= 0 || n_points < 0 || (n_points % n_time) != 0)
    return fail(CTX_E_INVALID, "n_points must be a multiple of n_time (points grouped per measurement)");
  if (n_chains <= 0) return CTX_OK;
  if (ctx_device_count() <= 0)
    return fail(CTX_E_NOGPU, "no CUDA device visible: catalax_b200 has no CPU fallback");
  cudaKernel_t k;
  int rc = get_kernel(m, CTX_TSIT5, 1, "ctx_rate_loglik", &k);
  if (rc) return rc;
  const unsigned block = 128;
  const unsigned long long grid = ((unsigned long long)n_chains * 32 + block - 1) / block;
  if (m->desc.dtype == CTX_F64) {
    RateLlArgs<double> a{(const double*)t, (const double*)y, (const double*)theta,
                         (const double*)consts, (const double*)inits, (const double*)data, mask,
                         (const double*)jac2, (const double*)extra2, (const double*)sigma,
                         (double*)out, n_chains, n_points, n_time, likelihood};
    void* params[] = {&a};
    CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(block), params, 0,
                              (cudaStream_t)stream));
  } else {
    RateLlArgs<float> a{(const float*)t, (const float*)y, (const float*)theta,
                        (const float*)consts, (const float*)inits, (const float*)data, mask,
                        (const float*)jac2, (const float*)extra2, (const float*)sigma,
                        (float*)out, n_chains, n_points, n_time, likelihood};
    void* params[] = {&a};
    CTX_CUDA(cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(block), params, 0,
                              (cudaStream_t)stream));
  }
  g_launches++;
  return CTX_OK;
}

size_t ctx_mlp_adjoint_workspace_bytes(ctx_model* m, const ctx_adjoint_cfg* cfg) {
  if (!m || !cfg || ctx_device_count() <= 0) return 0;
  AdjPlan p;
  if (adj_plan(m, cfg, p)) return 0;
  return p.total;
}

int ctx_mlp_adjoint_forward(ctx_model* m, const ctx_adjoint_cfg* cfg, const void* y0,
                            const void* theta, const void* ts, void* ys, int32_t* status,
                            int32_t* n_accepted, int32_t* n_rejected, void* workspace,
                            size_t ws_bytes, void* stream) {
  if (!m || !cfg || !ys) return fail(CTX_E_INVALID, "null argument");
  if (ctx_device_count() <= 0)
    return fail(CTX_E_NOGPU, "no CUDA device visible: catalax_b200 has no CPU fallback");
  if (m->desc.dtype == CTX_F64)
    return launch_adjoint<double>(m, cfg, false, y0, theta, ts, ys, nullptr, status, n_accepted,
                                  n_rejected, nullptr, nullptr, workspace, ws_bytes,
                                  (cudaStream_t)stream);
  return launch_adjoint<float>(m, cfg, false, y0, theta, ts, ys, nullptr, status, n_accepted,
                               n_rejected, nullptr, nullptr, workspace, ws_bytes,
                               (cudaStream_t)stream);
}

int ctx_mlp_adjoint_backward(ctx_model* m, const ctx_adjoint_cfg* cfg, const void* y0,
                             const void* theta, const void* ts, const void* gys, int32_t* status,
                             void* grad_theta, void* grad_y0, void* workspace, size_t ws_bytes,
                             void* stream) {
  if (!m || !cfg || !gys || !grad_theta || !status) return fail(CTX_E_INVALID, "null argument");
  if (ctx_device_count() <= 0)
    return fail(CTX_E_NOGPU, "no CUDA device visible: catalax_b200 has no CPU fallback");
  if (m->desc.dtype == CTX_F64)
    return launch_adjoint<double>(m, cfg, true, y0, theta, ts, nullptr, gys, status, nullptr,
                                  nullptr, grad_theta, grad_y0, workspace, ws_bytes,
                                  (cudaStream_t)stream);
  return launch_adjoint<float>(m, cfg, true, y0, theta, ts, nullptr, gys, status, nullptr,
                               nullptr, grad_theta, grad_y0, workspace, ws_bytes,
                               (cudaStream_t)stream);
}

// (re)allocate the staging area and the private streams of the host-buffer entry points;
// call with m->host_mu held
static int host_stage(ctx_model* m, size_t total) {
  int dev = 0;
  CTX_CUDA(cudaGetDevice(&dev));
  if (m->stage_device != dev) {
    // first call, or the caller switched devices: everything below belongs to one device
    if (m->stage) cudaFree(m->stage);
    m->stage = nullptr;
    m->stage_bytes = 0;
    for (int i = 0; i < 2; ++i) {
      if (m->ev_solved[i]) cudaEventDestroy(m->ev_solved[i]);
      if (m->ev_copied[i]) cudaEventDestroy(m->ev_copied[i]);
      m->ev_solved[i] = m->ev_copied[i] = nullptr;
    }
    if (m->s_compute) cudaStreamDestroy(m->s_compute);
    if (m->s_copy) cudaStreamDestroy(m->s_copy);
    m->s_compute = m->s_copy = nullptr;
    CTX_CUDA(cudaStreamCreateWithFlags(&m->s_compute, cudaStreamNonBlocking));
    CTX_CUDA(cudaStreamCreateWithFlags(&m->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CTX_CUDA(cudaEventCreateWithFlags(&m->ev_solved[i], cudaEventDisableTiming));
      CTX_CUDA(cudaEventCreateWithFlags(&m->ev_copied[i], cudaEventDisableTiming));
    }
    m->stage_device = dev;
  }
  if (m->stage_bytes < total) {
    if (m->stage) cudaFree(m->stage);
    m->stage = nullptr;
    m->stage_bytes = 0;
    CTX_CUDA(cudaMalloc(&m->stage, total));
    m->stage_bytes = total;
  }
  return CTX_OK;
}

// Rows per chunk of the host-buffer solve: ~512 MB of trajectory output per chunk, never fewer
// than 65536 rows (the warp-specialised kernel's threshold), a multiple of 4096; the whole batch
// when it is smaller than two chunks.  CTX_B200_HOST_CHUNK_ROWS overrides (0 = one chunk).
static size_t host_chunk_rows(size_t B, size_t row_bytes) {
  if (const char* e = getenv("CTX_B200_HOST_CHUNK_ROWS")) {
    const long long v = atoll(e);
    if (v <= 0) return B;
    return std::min<size_t>(B, (size_t)v);
  }
  size_t rows = (512ull << 20) / std::max<size_t>(row_bytes, 1);
  rows = std::max<size_t>(rows, 65536);
  rows = (rows + 4095) / 4096 * 4096;
  if (2 * rows > B) return B;
  return rows;
}

int ctx_solve_host(ctx_model* m, const ctx_solve_cfg* cfg, const void* y0, const void* theta,
                   const void* consts, const void* ts, void* ys, int32_t* status, int32_t* nacc,
                   int32_t* nrej) {
  if (!m || !cfg) return fail(CTX_E_INVALID, "null model or config");
  if (ctx_device_count() <= 0)
    return fail(CTX_E_NOGPU, "no CUDA device visible: catalax_b200 has no CPU fallback");
  if (cfg->inner > 0) return fail(CTX_E_INVALID, "ctx_solve_host: two-level batches take device buffers (ctx_solve)");
  if (cfg->batch <= 0) return CTX_OK;
  if (cfg->n_times <= 0) return fail(CTX_E_INVALID, "n_times must be >= 1");
  if (!y0 || !ts || !ys) return fail(CTX_E_INVALID, "null y0 / ts / ys");
  const ctx_model_desc& d = m->desc;
  const size_t w = d.dtype == CTX_F64 ? 8 : 4;
  const size_t B = (size_t)cfg->batch, T = (size_t)cfg->n_times, S = (size_t)d.n_states;
  const size_t row_out = T * S * w;
  const size_t R = host_chunk_rows(B, row_out);          // rows per chunk
  const size_t n_chunks = (B + R - 1) / R;
  const size_t n_buf = n_chunks > 1 ? 2 : 1;
  auto rows = [&](int ax, size_t n) { return cfg->in_axes[ax] == 0 ? n : (size_t)1; };
  auto al = [](size_t n) { return (n + 255) / 256 * 256; };
  ctx_solve_cfg ccfg = *cfg;
  ccfg.batch = (int64_t)R;
  ccfg.order = nullptr;          // (a DEVICE pointer indexing the whole batch: not applicable to chunks)
  // per buffer: inputs of one chunk, its outputs, counters, workspace
  const size_t n_y0 = al(rows(0, R) * S * w), n_th = al(rows(1, R) * d.n_params * w),
               n_c = al(rows(2, R) * d.n_consts * w), n_ts = al(rows(3, R) * T * w),
               n_ys = al(R * row_out), n_i = al(R * 4);
  const size_t n_ws = al(ctx_workspace_bytes(m, &ccfg));
  const size_t per_buf = n_y0 + n_th + n_c + n_ts + n_ys + 3 * n_i + n_ws;
  std::lock_guard<std::mutex> lk(m->host_mu);
  int rc = host_stage(m, per_buf * n_buf);
  if (rc) return rc;
  cudaStream_t sc = m->s_compute, sd = m->s_copy;
  // Pipeline: chunk k is solved on the compute stream into buffer k % 2 while the copy stream
  // drains chunk k - 1 to the host; a buffer is reused once the copy of chunk k - 2 has finished.
  for (size_t k = 0; k < n_chunks; ++k) {
    const size_t r0 = k * R, nr = std::min(R, B - r0);
    const int bi = (int)(k % n_buf);
    char* p = (char*)m->stage + per_buf * bi;
    char* dy0 = p; p += n_y0;
    char* dth = p; p += n_th;
    char* dc = p; p += n_c;
    char* dts = p; p += n_ts;
    char* dys = p; p += n_ys;
    int32_t* dst = (int32_t*)p; p += n_i;
    int32_t* dna = (int32_t*)p; p += n_i;
    int32_t* dnr = (int32_t*)p; p += n_i;
    char* dws = p;
    if (k >= n_buf) CTX_CUDA(cudaStreamWaitEvent(sc, m->ev_copied[bi], 0));
    auto off = [&](const void* base, int ax, size_t width) {
      return (const char*)base + (cfg->in_axes[ax] == 0 ? r0 * width * w : 0);
    };
    CTX_CUDA(cudaMemcpyAsync(dy0, off(y0, 0, S), rows(0, nr) * S * w, cudaMemcpyHostToDevice, sc));
    if (d.n_params)
      CTX_CUDA(cudaMemcpyAsync(dth, off(theta, 1, d.n_params), rows(1, nr) * d.n_params * w,
                               cudaMemcpyHostToDevice, sc));
    if (d.n_consts)
      CTX_CUDA(cudaMemcpyAsync(dc, off(consts, 2, d.n_consts), rows(2, nr) * d.n_consts * w,
                               cudaMemcpyHostToDevice, sc));
    if (cfg->in_axes[3] == 0 || k < n_buf)    // a shared grid is staged once per buffer
      CTX_CUDA(cudaMemcpyAsync(dts, off(ts, 3, T), rows(3, nr) * T * w, cudaMemcpyHostToDevice, sc));
    ccfg.batch = (int64_t)nr;
    rc = ctx_solve(m, &ccfg, dy0, dth, dc, dts, dys, dst, dna, dnr, dws, n_ws, sc);
    if (rc) { cudaStreamSynchronize(sc); cudaStreamSynchronize(sd); return rc; }
    CTX_CUDA(cudaEventRecord(m->ev_solved[bi], sc));
    CTX_CUDA(cudaStreamWaitEvent(sd, m->ev_solved[bi], 0));
    CTX_CUDA(cudaMemcpyAsync((char*)ys + r0 * row_out, dys, nr * row_out, cudaMemcpyDeviceToHost, sd));
    if (status) CTX_CUDA(cudaMemcpyAsync(status + r0, dst, nr * 4, cudaMemcpyDeviceToHost, sd));
    if (nacc) CTX_CUDA(cudaMemcpyAsync(nacc + r0, dna, nr * 4, cudaMemcpyDeviceToHost, sd));
    if (nrej) CTX_CUDA(cudaMemcpyAsync(nrej + r0, dnr, nr * 4, cudaMemcpyDeviceToHost, sd));
    CTX_CUDA(cudaEventRecord(m->ev_copied[bi], sd));
  }
  CTX_CUDA(cudaStreamSynchronize(sc));
  CTX_CUDA(cudaStreamSynchronize(sd));
  return CTX_OK;
}

int ctx_loglik_grad_host(ctx_model* m, const ctx_loglik_cfg* cfg, const void* y0s,
                         const void* theta, const void* consts, const void* ts, const void* data,
                         const uint8_t* mask, const void* sigma, void* out, void* partial,
                         int32_t* status) {
  if (!m || !cfg || !out) return fail(CTX_E_INVALID, "null argument");
  if (ctx_device_count() <= 0)
    return fail(CTX_E_NOGPU, "no CUDA device visible: catalax_b200 has no CPU fallback");
  if (cfg->n_chains <= 0 || cfg->n_series <= 0) return CTX_OK;
  if (cfg->n_times <= 0) return fail(CTX_E_INVALID, "n_times must be >= 1");
  const ctx_model_desc& d = m->desc;
  const size_t w = d.dtype == CTX_F64 ? 8 : 4;
  const size_t CH = (size_t)cfg->n_chains, M = (size_t)cfg->n_series, T = (size_t)cfg->n_times,
               S = (size_t)d.n_states, P = (size_t)d.n_params, Cn = (size_t)d.n_consts;
  const size_t KR = 1 + (size_t)d.n_theta_dirs + S;
  const size_t KP = KR + (size_t)(d.n_dirs - d.n_theta_dirs);
  auto al = [](size_t n) { return (n + 255) / 256 * 256; };
  const size_t n_y0 = al(M * S * w), n_th = al(CH * P * w), n_c = al(M * Cn * w),
               n_ts = al(M * T * w), n_d = al(M * T * S * w), n_m = al(M * T * S),
               n_sg = al(CH * S * w), n_out = al(CH * KR * w), n_part = al(CH * M * KP * w),
               n_st = al(CH * M * 4), n_ws = al(ctx_loglik_workspace_bytes(m, cfg));
  const size_t total = n_y0 + n_th + n_c + n_ts + n_d + n_m + n_sg + n_out + n_part + n_st + n_ws;
  ctx_loglik_cfg lcfg = *cfg;
  lcfg.order = nullptr;
  cfg = &lcfg;
  std::lock_guard<std::mutex> lk(m->host_mu);
  int rc = host_stage(m, total);
  if (rc) return rc;
  cudaStream_t s = m->s_compute;
  char* p = (char*)m->stage;
  char* dy0 = p; p += n_y0;
  char* dth = p; p += n_th;
  char* dc = p; p += n_c;
  char* dts = p; p += n_ts;
  char* dd = p; p += n_d;
  uint8_t* dm = (uint8_t*)p; p += n_m;
  char* dsg = p; p += n_sg;
  char* dout = p; p += n_out;
  char* dpart = p; p += n_part;
  int32_t* dst = (int32_t*)p; p += n_st;
  char* dws = p;
  CTX_CUDA(cudaMemcpyAsync(dy0, y0s, M * S * w, cudaMemcpyHostToDevice, s));
  CTX_CUDA(cudaMemcpyAsync(dth, theta, CH * P * w, cudaMemcpyHostToDevice, s));
  if (Cn) CTX_CUDA(cudaMemcpyAsync(dc, consts, M * Cn * w, cudaMemcpyHostToDevice, s));
  CTX_CUDA(cudaMemcpyAsync(dts, ts, M * T * w, cudaMemcpyHostToDevice, s));
  CTX_CUDA(cudaMemcpyAsync(dd, data, M * T * S * w, cudaMemcpyHostToDevice, s));
  if (mask) CTX_CUDA(cudaMemcpyAsync(dm, mask, M * T * S, cudaMemcpyHostToDevice, s));
  CTX_CUDA(cudaMemcpyAsync(dsg, sigma, CH * S * w, cudaMemcpyHostToDevice, s));
  rc = ctx_loglik_grad(m, cfg, dy0, dth, dc, dts, dd, mask ? dm : nullptr, dsg, dout, dpart, dst,
                       dws, n_ws, s);
  if (rc) { cudaStreamSynchronize(s); return rc; }
  CTX_CUDA(cudaMemcpyAsync(out, dout, CH * KR * w, cudaMemcpyDeviceToHost, s));
  if (partial) CTX_CUDA(cudaMemcpyAsync(partial, dpart, CH * M * KP * w, cudaMemcpyDeviceToHost, s));
  if (status) CTX_CUDA(cudaMemcpyAsync(status, dst, CH * M * 4, cudaMemcpyDeviceToHost, s));
  CTX_CUDA(cudaStreamSynchronize(s));
  return CTX_OK;
}

#ifndef CTX_BUILD_STAMP
#define CTX_BUILD_STAMP "unstamped"
#endif
// hash of the sources this library was built from (catalax_b200/_build.py compares it with the
// tree at load time, so a stale library next to newer sources is rebuilt or refused)
const char* ctx_build_stamp(void) {
  static const char marked[] = "CTXB200-SOURCE-STAMP:" CTX_BUILD_STAMP;
  return marked + sizeof("CTXB200-SOURCE-STAMP:") - 1;
}

}  // extern "C"
