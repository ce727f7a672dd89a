#!/bin/bash
# round 2, GPU call C: new parity / caller tests + v2 layout sweep + v1 occupancy sweep
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_solve.py tests/test_gpu_reference_simulation.py tests/test_gpu_reference_pins.py tests/test_gpu_posterior_callers.py tests/test_gpu_edge_cases.py -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|residual|fitted|v2 x64|row [01]:" gpurun_out/r2c_pytest.log | tail -20
export CTX_B200_NO_DISK_CACHE=1
{
for T in 16 64 128 250 1000; do
  echo "== T=$T v1 (kernel 2) default"; python scripts/dev_bench.py --T $T --dtype float32 --kernel 2 --iters 4 2>&1 | head -1 | cut -c1-170
  for fl in "" "-DV2_NP=10 -DV2_PPC=5 -DV2_SLOTS=2" "-DV2_NP=9 -DV2_PPC=3 -DV2_SLOTS=2" "-DV2_NP=8 -DV2_PPC=2 -DV2_SLOTS=3" "-DV2_NP=8 -DV2_PPC=4 -DV2_SLOTS=3"; do
    echo "== T=$T v2 flags='$fl'"; CTX_B200_NVRTC_FLAGS="$fl" timeout 300 python scripts/dev_bench.py --T $T --dtype float32 --kernel 5 --iters 4 2>&1 | head -1 | cut -c1-170
  done
done
for mb in 5 6; do
  echo "== T=16 v1 MIN_BLOCKS=$mb"; CTX_B200_NVRTC_FLAGS="-DCTX_MIN_BLOCKS=$mb" python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 4 2>&1 | head -1 | cut -c1-170
done
echo "== f64 T=16/1000 v2 layouts"
for fl in "" "-DV2_NP=12 -DV2_PPC=3 -DV2_SLOTS=2"; do
  echo "== f64 flags='$fl'"; CTX_B200_NVRTC_FLAGS="$fl" timeout 300 python scripts/dev_bench.py --T 16 1000 --dtype float64 --kernel 5 --iters 3 2>&1 | head -2 | cut -c1-170
done
} > gpurun_out/r2c_sweep.log 2>&1
cat gpurun_out/r2c_sweep.log
