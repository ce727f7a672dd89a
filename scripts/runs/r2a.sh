#!/bin/bash
# round 2, GPU call A: baseline numbers + A/B that need no code change
set -u
mkdir -p gpurun_out
export CTX_B200_NO_DISK_CACHE=1
echo "== baseline f32 T16/T1000 k2 k5" > gpurun_out/r2a_ab.log
python scripts/dev_bench.py --T 16 1000 --dtype float32 --kernel 2 5 --iters 5 2>&1 | head -4 | cut -c1-220 >> gpurun_out/r2a_ab.log
echo "== prec-div=false" >> gpurun_out/r2a_ab.log
CTX_B200_NVRTC_FLAGS="--prec-div=false" python scripts/dev_bench.py --T 16 1000 --dtype float32 --kernel 2 5 --iters 5 2>&1 | head -4 | cut -c1-220 >> gpurun_out/r2a_ab.log
echo "== prec-div=false prec-sqrt=false ftz" >> gpurun_out/r2a_ab.log
CTX_B200_NVRTC_FLAGS="--prec-div=false --prec-sqrt=false" python scripts/dev_bench.py --T 16 1000 --dtype float32 --kernel 2 5 --iters 5 2>&1 | head -4 | cut -c1-220 >> gpurun_out/r2a_ab.log
echo "== parity f32 with prec-div=false" >> gpurun_out/r2a_ab.log
CTX_B200_NVRTC_FLAGS="--prec-div=false" python -m pytest tests/test_gpu_parity_solve.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -15 >> gpurun_out/r2a_ab.log
echo "== pcie" >> gpurun_out/r2a_ab.log
PCIE_GB=12.6 python scripts/pcie_ceiling.py >> gpurun_out/r2a_ab.log 2>&1
echo "== config3 stability" >> gpurun_out/r2a_ab.log
python scripts/dev_bench3_stab.py >> gpurun_out/r2a_ab.log 2>&1
cat gpurun_out/r2a_ab.log
