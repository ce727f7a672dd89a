#!/bin/bash
# round 2, GPU call D: fixed tests + XLA shim + composition goldens; dynamic-consumer sweep
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_solve.py tests/test_gpu_reference_simulation.py tests/test_gpu_posterior_callers.py tests/test_gpu_xla_shim.py tests/test_neural_composition_golden.py tests/test_eqx_checkpoint.py tests/test_gpu_parity_neural.py -m gpu -q -s > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|v2 x64|row [01]:" gpurun_out/r2d_pytest.log | tail -20
export CTX_B200_NO_DISK_CACHE=1
{
for T in 1000 250 16; do
  for fl in "" "-DV2_NP=6 -DV2_DYN=6" "-DV2_NP=5 -DV2_DYN=7" "-DV2_NP=4 -DV2_DYN=8" "-DV2_NP=6 -DV2_DYN=6 -DV2_SLOTS=3" "-DV2_NP=7 -DV2_DYN=5" "-DV2_NP=8 -DV2_DYN=4 -DV2_SLOTS=3" "-DV2_NP=9 -DV2_DYN=3 -DV2_SLOTS=2"; do
    echo "== T=$T v2 flags='$fl'"; CTX_B200_NVRTC_FLAGS="$fl" timeout 300 python scripts/dev_bench.py --T $T --dtype float32 --kernel 5 --iters 5 2>&1 | head -1 | cut -c1-170
  done
done
echo "== f64 T=1000"
for fl in "" "-DV2_NP=8 -DV2_DYN=8" "-DV2_NP=6 -DV2_DYN=10"; do
  echo "== f64 flags='$fl'"; CTX_B200_NVRTC_FLAGS="$fl" timeout 300 python scripts/dev_bench.py --T 1000 --dtype float64 --kernel 5 --iters 3 2>&1 | head -1 | cut -c1-170
done
} > gpurun_out/r2d_sweep.log 2>&1
cat gpurun_out/r2d_sweep.log
