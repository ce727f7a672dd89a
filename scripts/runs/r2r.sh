#!/bin/bash
# round 2, GPU call R (1 GPU): which part of --use_fast_math makes the solve faster (T=16: 1.45 -> 0.99 ms)?
set -u
mkdir -p gpurun_out
{
for flags in "" "--use_fast_math" "--use_fast_math --prec-div=true" "--use_fast_math --ftz=false" "--use_fast_math --prec-sqrt=true" "--prec-div=false" "--ftz=true" "--prec-div=false --ftz=true"; do
echo "== NVRTC flags: [$flags]"; CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench.py --kernel 0 --dtype float32 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k: d[k] for k in ('T', 'ms', 'steps', 'tail')})
"
done
} > gpurun_out/r2r_fastmath.log 2>&1
cat gpurun_out/r2r_fastmath.log
