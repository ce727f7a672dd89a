#!/bin/bash
# round 2, GPU call T (1 GPU): bit-exact zero-numerator guard of the IEEE division vs the SFU division
set -u
mkdir -p gpurun_out
{
for flags in "" "-DCTX_DIV_GUARD=0" "--ftz=true" "-DCTX_FAST_DIV=3" "-DCTX_FAST_DIV=3 --prec-sqrt=false --ftz=true"; do
echo "== f32, NVRTC flags: [$flags]"; CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench.py --kernel 0 --dtype float32 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k: d[k] for k in ('T', 'ms', 'steps', 'tail')})
"
done
for flags in "" "-DCTX_DIV_GUARD=0"; do
echo "== f64, NVRTC flags: [$flags]"; CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench.py --kernel 0 --dtype float64 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k: d[k] for k in ('T', 'ms', 'steps', 'tail')})
"
echo "== K2 f32 [$flags]"; CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench_nuts.py f32 2>&1 | tail -2 | cut -c1-300
done
} > gpurun_out/r2t_div.log 2>&1
cat gpurun_out/r2t_div.log
