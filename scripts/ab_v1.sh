#!/bin/bash
# A/B matrix of ctx_solve_v1 build switches (developer tool, run under gpurun)
# usage: ab_v1.sh "flags|ENV=..|dtype" ...
export CTX_B200_NO_DISK_CACHE=1
run() { echo "== $1 | $2 | $3"; env CTX_B200_NVRTC_FLAGS="$1" $2 python scripts/dev_bench.py --T 16 1000 --dtype ${3:-float32} --kernel 2 2>&1 | head -2 | cut -c1-150; }
for cfg in "$@"; do
  IFS='|' read -r flags envs dt <<< "$cfg"
  run "$flags" "$envs" "$dt"
done
