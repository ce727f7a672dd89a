#!/bin/bash
# round 2, GPU call F (1 GPU): direct-save variant at small T, K2 occupancy sweep, new kernels' tests
set -u
mkdir -p gpurun_out
export CTX_B200_NO_DISK_CACHE=1
{
for T in 16 32 64 128; do
  for ds in 0 1; do
    echo "== v1 T=$T DIRECT_SAVE=$ds"; CTX_B200_DIRECT_SAVE=$ds python scripts/dev_bench.py --T $T --dtype float32 --kernel 2 --iters 5 2>&1 | head -1 | cut -c1-170
  done
done
echo "== v1 T=16 DIRECT_SAVE=1 MIN_BLOCKS=6"; CTX_B200_DIRECT_SAVE=1 CTX_B200_NVRTC_FLAGS="-DCTX_MIN_BLOCKS=6" python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 5 2>&1 | head -1 | cut -c1-170
echo "== v1 T=16 DIRECT_SAVE=1 WARPS=8"; CTX_B200_DIRECT_SAVE=1 CTX_B200_NVRTC_FLAGS="-DCTX_MIN_BLOCKS=3" python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 5 2>&1 | head -1 | cut -c1-170
for T in 16 32; do
  for ds in 0 1; do
    echo "== f64 v1 T=$T DIRECT_SAVE=$ds"; CTX_B200_DIRECT_SAVE=$ds python scripts/dev_bench.py --T $T --dtype float64 --kernel 2 --iters 3 2>&1 | head -1 | cut -c1-170
  done
done
echo "== K2 occupancy"
for mb in "" "-DCTX_MIN_BLOCKS=3" "-DCTX_MIN_BLOCKS=4" "-DCTX_MIN_BLOCKS=5"; do
  echo "== K2 f32 flags='$mb'"; CTX_B200_NVRTC_FLAGS="$mb" python scripts/dev_bench_nuts.py f32 2>&1 | tail -1
done
for mb in "" "-DCTX_MIN_BLOCKS=2" "-DCTX_MIN_BLOCKS=3"; do
  echo "== K2 f64 flags='$mb'"; CTX_B200_NVRTC_FLAGS="$mb" python scripts/dev_bench_nuts.py f64 2>&1 | tail -1
done
} > gpurun_out/r2f_sweep.log 2>&1
cat gpurun_out/r2f_sweep.log
unset CTX_B200_NO_DISK_CACHE
python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity_solve.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
