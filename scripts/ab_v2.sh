#!/bin/bash
# A/B of the warp-specialised kernel's build switches (developer tool, run under gpurun)
export CTX_B200_NO_DISK_CACHE=1
for cfg in "$@"; do
  IFS='|' read -r dt flags <<< "$cfg"
  echo "== $dt | $flags"
  env CTX_B200_NVRTC_FLAGS="$flags" timeout 200 python scripts/dev_v2.py $dt 2>&1 | tail -3 | cut -c1-160
done
