// ctx_p2p.cu -- one-shot all-reduce(sum) over NVLink peer memory for the path's only collective.
//
// Series-sharded NUTS (reference: the sum over measurement series inside the numpyro
// model, catalax/mcmc/mcmc.py:486-491, when the series of one posterior live on several GPUs):
// every rank holds a [chains, 1+P+n_obs] block of partial log-likelihoods / gradients that must be
// summed over the ranks once per leapfrog step.  At 16 384 chains that is 384 KiB: a latency-bound
// message for which a ring / tree collective pays several launch + hop latencies.  Here every
// rank's block lives in a buffer that all ranks of the node have mapped (cudaIpc over NVLink /
// NVSwitch) and ONE kernel per evaluation does everything:
//   phase 0  copy the rank's block into its own symmetric buffer (slot = epoch & 1);
//   phase 1  the last block to finish that copy raises this rank's flag in every peer's flag array
//            (st.release.sys after a system-scope fence);
//   phase 2  every block waits until all ranks' flags carry the current epoch (ld.acquire.sys);
//   phase 3  every thread sums its elements over the peers' buffers IN RANK ORDER (P2P loads), so
//            all ranks get bit-identical results; the last block to finish bumps the epoch.
// The epoch and both completion counters live in device memory, so the kernel is replayable from a
// CUDA graph (no host-side argument changes between evaluations).  Two slots make the protocol
// safe without a trailing barrier: a rank overwrites slot s again only two evaluations later, after
// it has seen every peer's flag for the evaluation in between, which the peer raises only once its
// reads of slot s have completed (stream order).  All spins are bounded and trap instead of hanging.
#include "../../include/ctx_b200.h"

#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstring>
#include <string>

namespace {

constexpr int kMaxWorld = 16;
constexpr int kBlock = 256;

struct Ctl {                  // device-resident control block (local to the rank)
  unsigned epoch;             // evaluations completed so far
  unsigned arrived;           // blocks that finished phase 0 of the current evaluation
  unsigned finished;          // blocks that finished phase 3
  unsigned pad;
};

struct Peers {                // device-resident pointer table
  void* buf[kMaxWorld][2];    // peer r's slot buffers
  unsigned* flags[kMaxWorld]; // peer r's flag array: [2 slots][world]
};

thread_local std::string p2p_err;
int p2p_fail(int code, const std::string& msg) { p2p_err = msg; return code; }

template <typename real>
__global__ void __launch_bounds__(kBlock)
ctx_p2p_allreduce_k(const Peers* __restrict__ peers, Ctl* ctl, real* inout, const size_t n,
                    const int rank, const int world) {
  const unsigned ep = *reinterpret_cast<volatile unsigned*>(&ctl->epoch);
  const int slot = (int)(ep & 1u);
  const unsigned want = ep + 1u;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // ---- phase 0: my block -> my symmetric buffer
  real* mine = reinterpret_cast<real*>(peers->buf[rank][slot]);
  for (size_t i = tid; i < n; i += stride) mine[i] = inout[i];
  __threadfence_system();
  __syncthreads();
  // ---- phase 1: the last block raises my flag at every peer
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(&ctl->arrived, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x < world) {
    __threadfence_system();
    unsigned* f = peers->flags[threadIdx.x] + slot * world + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(want) : "memory");
  }
  // ---- phase 2: wait for every rank's flag of this evaluation
  if (threadIdx.x < world) {
    const unsigned* f = peers->flags[rank] + slot * world + threadIdx.x;
    unsigned v = 0, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v >= want) break;
      if (++spins > (1u << 27)) { printf("ctx_p2p: rank %d timed out waiting for rank %d\n", rank, (int)threadIdx.x); __trap(); }
      __nanosleep(40);
    } while (true);
  }
  __syncthreads();
  // ---- phase 3: sum over the peers' buffers in rank order
  for (size_t i = tid; i < n; i += stride) {
    real s = reinterpret_cast<const real*>(peers->buf[0][slot])[i];
    for (int r = 1; r < world; ++r) s += reinterpret_cast<const real*>(peers->buf[r][slot])[i];
    inout[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl->finished, 1u) == gridDim.x - 1) {   // every block has read `epoch`
      ctl->arrived = 0u;
      ctl->finished = 0u;
      __threadfence();
      ctl->epoch = want;
    }
  }
}

}  // namespace

struct ctx_p2p {
  int rank = 0, world = 1, device = 0;
  size_t bytes = 0;            // per slot
  char* base = nullptr;        // [flags 256 B][slot 0][slot 1]
  size_t slot_off = 256;
  void* peer_base[kMaxWorld] = {};
  Peers* d_peers = nullptr;
  Ctl* d_ctl = nullptr;
  bool connected = false;
};

extern "C" {

size_t ctx_p2p_last_error(char* buf, size_t n) {
  if (buf && n) {
    const size_t k = p2p_err.size() < n - 1 ? p2p_err.size() : n - 1;
    memcpy(buf, p2p_err.data(), k);
    buf[k] = '\0';
  }
  return p2p_err.size();
}

int ctx_p2p_create(int32_t rank, int32_t world, size_t bytes, ctx_p2p** out, void* handle_out) {
  if (!out || !handle_out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || bytes == 0)
    return p2p_fail(CTX_E_INVALID, "ctx_p2p_create: bad argument (world <= 16)");
  ctx_p2p* p = new ctx_p2p;
  p->rank = rank; p->world = world;
  p->bytes = (bytes + 255) / 256 * 256;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->base, p->slot_off + 2 * p->bytes);
  if (e == cudaSuccess) e = cudaMemset(p->base, 0, p->slot_off + 2 * p->bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_peers, sizeof(Peers));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_ctl, sizeof(Ctl));
  if (e == cudaSuccess) e = cudaMemset(p->d_ctl, 0, sizeof(Ctl));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->base);
  if (e != cudaSuccess) {
    ctx_p2p_free(p);
    return p2p_fail(e == cudaErrorNoDevice ? CTX_E_NOGPU : CTX_E_CUDA, std::string("ctx_p2p_create: ") + cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  memcpy(handle_out, &h, 64);
  *out = p;
  return CTX_OK;
}

int ctx_p2p_connect(ctx_p2p* p, const void* all_handles) {
  if (!p || !all_handles) return p2p_fail(CTX_E_INVALID, "ctx_p2p_connect: null argument");
  Peers host;
  memset(&host, 0, sizeof host);
  for (int r = 0; r < p->world; ++r) {
    void* base = p->base;
    if (r != p->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)all_handles + 64 * r, 64);
      cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return p2p_fail(CTX_E_CUDA, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
      p->peer_base[r] = base;
    }
    host.flags[r] = reinterpret_cast<unsigned*>(base);
    host.buf[r][0] = (char*)base + p->slot_off;
    host.buf[r][1] = (char*)base + p->slot_off + p->bytes;
  }
  cudaError_t e = cudaMemcpy(p->d_peers, &host, sizeof host, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return p2p_fail(CTX_E_CUDA, std::string("ctx_p2p_connect: ") + cudaGetErrorString(e));
  p->connected = true;
  return CTX_OK;
}

int ctx_p2p_allreduce(ctx_p2p* p, int32_t dtype, size_t n, void* inout, void* stream) {
  if (!p || !inout || !p->connected) return p2p_fail(CTX_E_INVALID, "ctx_p2p_allreduce: not connected");
  const size_t w = dtype == CTX_F64 ? 8 : 4;
  if (n * w > p->bytes) return p2p_fail(CTX_E_INVALID, "ctx_p2p_allreduce: message larger than the symmetric buffer");
  if (n == 0) return CTX_OK;
  // every block must be resident at once (blocks wait for each other's copies across ranks)
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
  size_t grid = (n + kBlock - 1) / kBlock;
  if (grid > (size_t)sms * 4) grid = (size_t)sms * 4;
  cudaError_t e;
  if (dtype == CTX_F64)
    ctx_p2p_allreduce_k<double><<<(unsigned)grid, kBlock, 0, (cudaStream_t)stream>>>(p->d_peers, p->d_ctl, (double*)inout, n, p->rank, p->world);
  else
    ctx_p2p_allreduce_k<float><<<(unsigned)grid, kBlock, 0, (cudaStream_t)stream>>>(p->d_peers, p->d_ctl, (float*)inout, n, p->rank, p->world);
  e = cudaGetLastError();
  if (e != cudaSuccess) return p2p_fail(CTX_E_CUDA, std::string("ctx_p2p_allreduce: ") + cudaGetErrorString(e));
  return CTX_OK;
}

void ctx_p2p_free(ctx_p2p* p) {
  if (!p) return;
  for (int r = 0; r < p->world; ++r)
    if (p->peer_base[r]) cudaIpcCloseMemHandle(p->peer_base[r]);
  if (p->d_peers) cudaFree(p->d_peers);
  if (p->d_ctl) cudaFree(p->d_ctl);
  if (p->base) cudaFree(p->base);
  delete p;
}

}  // extern "C"
