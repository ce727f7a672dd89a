#!/bin/bash
# Under gpurun: full ncu capture (with source) of ctx_solve_v1 at config 2 / T=1000 / f32.
# NVRTC records the integrator header by its include name; copy it next to the cwd so that
# --import-source finds it.  usage: profile_v1.sh TAG ["extra nvrtc flags"]
set -u
TAG=${1:-r1c}
export CTX_B200_NVRTC_FLAGS="${2:-}"
mkdir -p gpurun_out
cp catalax_b200/csrc/ctx_integrator.cuh catalax_b200/csrc/ctx_mlp_tc.cuh .
python scripts/dev_bench.py --T 16 1000 --dtype float32 --kernel 2 > gpurun_out/${TAG}_dev_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctx_solve_v1 -s 6 -c 1 \
    -o gpurun_out/${TAG}_solve_full python scripts/dev_bench.py --T 16 1000 --dtype float32 --kernel 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
rm -f ctx_integrator.cuh ctx_mlp_tc.cuh
head -2 gpurun_out/${TAG}_dev_plain.log
