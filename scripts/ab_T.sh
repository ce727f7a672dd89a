#!/bin/bash
# A/B of ctx_solve_v1 build switches over several T (developer tool, run under gpurun)
export CTX_B200_NO_DISK_CACHE=1
for cfg in "$@"; do
  echo "== $cfg"
  env CTX_B200_NVRTC_FLAGS="$cfg" python scripts/dev_bench.py --T 16 64 250 1000 --dtype float32 --kernel 2 2>&1 | head -4 | cut -c1-90
done
