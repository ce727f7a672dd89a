#!/bin/bash
# A/B of ctx_solve_v1 build switches over several T (developer tool, run under gpurun)
# usage: ab_T.sh "<dtype>|<flags>" ...
export CTX_B200_NO_DISK_CACHE=1
for cfg in "$@"; do
  IFS='|' read -r dt flags <<< "$cfg"
  echo "== $dt | $flags"
  env CTX_B200_WIDE_SAVE=0 CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench.py --T 16 250 1000 --dtype $dt --kernel 2 2>&1 | head -3 | cut -c1-90
done
