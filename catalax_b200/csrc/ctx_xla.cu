// ctx_xla.cu -- header-free XLA GPU custom-call shims over the C ABI (include/ctx_b200.h).
//
// North star: "Python host code calls CUDA through a thin jax.ffi / XLA custom-call C-ABI layer."
// The reference's seam is the jitted closure of catalax/tools/simulation.py:160-196, called at
// catalax/model/model.py:699 and catalax/mcmc/mcmc.py:448 from inside jit-compiled JAX programs.
// For our kernels to run INSIDE such a program XLA must be able to call them on its own stream
// with its own buffers; the entry points below have exactly the signature of XLA's GPU
// custom-call ABI ("API_VERSION_ORIGINAL": stream, buffers[], opaque bytes), which needs no
// jaxlib / XLA header.  buffers[] holds the operands in order, then the results in order (a
// scratch workspace is declared as one more result); `opaque` is one of the packed descriptors
// of include/ctx_b200.h, built on the Python side (INTEGRATION.md section 3 shows the
// registration: jax.ffi.register_ffi_target(..., api_version=0) + jax.ffi.ffi_call).
//
// This ABI has no status channel, so a failing call records its code and message in a
// process-wide slot (XLA runs custom calls on its own threads, the thread-local message of
// ctx_last_error would be invisible to Python) that ctx_xla_last_status / ctx_xla_last_error
// return; the Python wrapper checks it after block_until_ready().
#include "../../include/ctx_b200.h"

#include <atomic>
#include <cstring>
#include <mutex>
#include <string>

namespace {
std::atomic<int> g_xla_status{0};
std::mutex g_xla_mu;
std::string g_xla_msg;

void xla_fail(int rc, const char* where) {
  char buf[4096];
  buf[0] = '\0';
  ctx_last_error(buf, sizeof buf);
  std::lock_guard<std::mutex> lk(g_xla_mu);
  g_xla_msg = std::string(where) + ": " + buf;
  g_xla_status.store(rc);
}

void xla_fail_msg(int rc, const char* msg) {
  std::lock_guard<std::mutex> lk(g_xla_mu);
  g_xla_msg = msg;
  g_xla_status.store(rc);
}

template <typename D>
bool unpack(const char* opaque, size_t n, D* d, const char* where) {
  if (!opaque || n != sizeof(D)) {
    xla_fail_msg(CTX_E_INVALID, (std::string(where) + ": opaque descriptor has the wrong size").c_str());
    return false;
  }
  memcpy(d, opaque, sizeof(D));   // XLA gives no alignment guarantee for the opaque bytes
  if (d->abi != CTX_ABI_VERSION || !d->model) {
    xla_fail_msg(CTX_E_INVALID, (std::string(where) + ": descriptor ABI version / model handle").c_str());
    return false;
  }
  return true;
}
}  // namespace

extern "C" {

size_t ctx_xla_desc_bytes(int32_t which) {
  switch (which) {
    case 0: return sizeof(ctx_xla_solve_desc);
    case 1: return sizeof(ctx_xla_loglik_desc);
    default: return 0;
  }
}

int ctx_xla_last_status(void) { return g_xla_status.exchange(0); }

size_t ctx_xla_last_error(char* buf, size_t n) {
  std::lock_guard<std::mutex> lk(g_xla_mu);
  if (buf && n) {
    const size_t k = g_xla_msg.size() < n - 1 ? g_xla_msg.size() : n - 1;
    memcpy(buf, g_xla_msg.data(), k);
    buf[k] = '\0';
  }
  return g_xla_msg.size();
}

// operands: y0, theta, consts, ts      results: ys, status, n_accepted, n_rejected, workspace
void ctx_xla_solve(void* stream, void** buffers, const char* opaque, size_t opaque_len) {
  ctx_xla_solve_desc d;
  if (!unpack(opaque, opaque_len, &d, "ctx_xla_solve")) return;
  ctx_model* m = reinterpret_cast<ctx_model*>(static_cast<uintptr_t>(d.model));
  const int rc = ctx_solve(m, &d.cfg, buffers[0], buffers[1], buffers[2], buffers[3], buffers[4],
                           (int32_t*)buffers[5], (int32_t*)buffers[6], (int32_t*)buffers[7],
                           buffers[8], (size_t)d.ws_bytes, stream);
  if (rc) xla_fail(rc, "ctx_xla_solve");
}

// operands: y0, theta, consts, ts      results: ys, sens, status, n_accepted, n_rejected, workspace
void ctx_xla_solve_sens(void* stream, void** buffers, const char* opaque, size_t opaque_len) {
  ctx_xla_solve_desc d;
  if (!unpack(opaque, opaque_len, &d, "ctx_xla_solve_sens")) return;
  ctx_model* m = reinterpret_cast<ctx_model*>(static_cast<uintptr_t>(d.model));
  const int rc = ctx_solve_sens(m, &d.cfg, buffers[0], buffers[1], buffers[2], buffers[3],
                                buffers[4], buffers[5], (int32_t*)buffers[6], (int32_t*)buffers[7],
                                (int32_t*)buffers[8], buffers[9], (size_t)d.ws_bytes, stream);
  if (rc) xla_fail(rc, "ctx_xla_solve_sens");
}

// operands: y0s, theta, consts, ts, data, mask (uint8; ignored unless desc.has_mask), sigma
// results:  out, partial, status, workspace
void ctx_xla_loglik_grad(void* stream, void** buffers, const char* opaque, size_t opaque_len) {
  ctx_xla_loglik_desc d;
  if (!unpack(opaque, opaque_len, &d, "ctx_xla_loglik_grad")) return;
  ctx_model* m = reinterpret_cast<ctx_model*>(static_cast<uintptr_t>(d.model));
  const int rc = ctx_loglik_grad(m, &d.cfg, buffers[0], buffers[1], buffers[2], buffers[3],
                                 buffers[4], d.has_mask ? (const uint8_t*)buffers[5] : nullptr,
                                 buffers[6], buffers[7], buffers[8], (int32_t*)buffers[9],
                                 buffers[10], (size_t)d.ws_bytes, stream);
  if (rc) xla_fail(rc, "ctx_xla_loglik_grad");
}

}  // extern "C"
