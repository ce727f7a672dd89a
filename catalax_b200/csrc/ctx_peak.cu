// ctx_peak.cu -- measured FFMA / DFMA peak of the device (roofline denominator for the
// compute-bound kernels: MEASURED_PEAKS.json only carries HBM and bf16-GEMM figures).
#include "../../include/ctx_b200.h"

#include <cuda_runtime_api.h>

namespace {

template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
  // 16 independent FMA chains per thread: enough ILP to saturate the pipe at full occupancy
  T x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = a + (T)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = x[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == (T)123456789) out[0] = s;  // never true: keeps the loop alive
}

template <typename T>
int measure(double* tflops) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return CTX_E_NOGPU;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  T* out = nullptr;
  if (cudaMalloc(&out, sizeof(T)) != cudaSuccess) return CTX_E_CUDA;
  const int iters = 4096, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    fma_peak_kernel<T><<<blocks, threads>>>(out, iters, (T)0.999, (T)0.001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16.0 * iters * (double)blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? CTX_OK : CTX_E_CUDA;
}

}  // namespace

extern "C" int ctx_measure_fma_peak(int32_t dtype, double* tflops) {
  if (!tflops) return CTX_E_INVALID;
  if (ctx_device_count() <= 0) return CTX_E_NOGPU;
  return dtype == CTX_F64 ? measure<double>(tflops) : measure<float>(tflops);
}
