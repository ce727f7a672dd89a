// ctx_sched.cu -- cost classes for the scheduling of one solve launch.
//
// One launch of the persistent solve kernels is bound by the critical path of its longest
// trajectories: rows are handed out in queue order and a 600-step row that is
// claimed when the queue runs dry finishes long after everything else.  A caller that solves
// similar batches repeatedly passes `order` (cost-ordered from the previous step counts).  For a
// one-shot solve ctx_solve can (opt-in, ctx_solve_cfg.reserved bit 1) first run a PROBE of the same
// batch -- the same integrator at 100 x rtol, capped at 32 step attempts, one requested time -- whose
// only product is a per-row bit "still running at the cap".  The kernels below turn that bit into
// the `order` permutation of the real launch: flagged rows first (index order among themselves;
// the strided queue start spreads them one per lane), every other row behind them in index order
// (a STABLE partition, so the bulk keeps its memory locality).  Three tiny launches, no host round
// trip; results of the solve do not depend on any of it.
// Measured (B200, config 2): NOT a win there -- the long rows of that workload are long only at
// the tight tolerance, a probe cheap enough to pay (rtol 1e-2, 0.18 ms) flags none of them and
// the launch stays at 3.53 ms + the probe; hence opt-in.
#include "../../include/ctx_b200.h"

#include <cuda_runtime_api.h>

namespace {

constexpr int kChunk = 1024;

__global__ void __launch_bounds__(kChunk)
sched_count(const int* __restrict__ nacc, const int* __restrict__ nrej, const int cap, const long long B,
            int* __restrict__ blk_cnt) {
  const long long b = (long long)blockIdx.x * kChunk + threadIdx.x;
  const int heavy = (b < B && nacc[b] + nrej[b] >= cap) ? 1 : 0;
  const int c = __syncthreads_count(heavy);
  if (threadIdx.x == 0) blk_cnt[blockIdx.x] = c;
}

// exclusive scan of the per-chunk counts by ONE block (<= 2^22 rows -> <= 4096 chunks)
__global__ void __launch_bounds__(1024) sched_scan(int* blk_cnt, const int n_blk) {
  __shared__ int part[1024];
  const int per = (n_blk + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(lo + per, n_blk);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += blk_cnt[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - s;      // exclusive prefix of this thread's range
  for (int i = lo; i < hi; ++i) { const int c = blk_cnt[i]; blk_cnt[i] = run; run += c; }
  if (threadIdx.x == 1023) blk_cnt[n_blk] = part[1023];   // total number of heavy rows
}

__global__ void __launch_bounds__(kChunk)
sched_place(const int* __restrict__ nacc, const int* __restrict__ nrej, const int cap, const long long B,
            const int* __restrict__ blk_off, const int n_blk, int* __restrict__ order) {
  __shared__ int warp_cnt[kChunk / 32];
  const long long b = (long long)blockIdx.x * kChunk + threadIdx.x;
  const int heavy = (b < B && nacc[b] + nrej[b] >= cap) ? 1 : 0;
  const unsigned m = __ballot_sync(0xffffffffu, heavy);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[wid] = __popc(m);
  __syncthreads();
  int before = __popc(m & ((1u << lane) - 1u));
  for (int i = 0; i < wid; ++i) before += warp_cnt[i];
  if (b < B) {
    const int heavy_before = blk_off[blockIdx.x] + before;      // heavy rows with a smaller index
    const int n_heavy = blk_off[n_blk];
    const long long pos = heavy ? (long long)heavy_before : (long long)n_heavy + (b - heavy_before);
    order[pos] = (int)b;
  }
}

}  // namespace

// order[B] (int32) <- stable partition of the rows by (n_accepted + n_rejected >= cap); scratch:
// (B + 1023) / 1024 + 1 ints.  B <= 2^22.  Returns a cudaError_t as int.
extern "C" int ctx_sched_heavy_first(const int* nacc, const int* nrej, int cap, long long B, int* order,
                                     int* scratch, void* stream) {
  const int n_blk = (int)((B + kChunk - 1) / kChunk);
  cudaStream_t s = (cudaStream_t)stream;
  sched_count<<<n_blk, kChunk, 0, s>>>(nacc, nrej, cap, B, scratch);
  sched_scan<<<1, 1024, 0, s>>>(scratch, n_blk);
  sched_place<<<n_blk, kChunk, 0, s>>>(nacc, nrej, cap, B, scratch, n_blk, order);
  return (int)cudaGetLastError();
}
