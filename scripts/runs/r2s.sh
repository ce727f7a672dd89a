#!/bin/bash
# round 2, GPU call S (1 GPU): where the IEEE divisions cost (const-state skip in the error norm; SFU division in the
# integrator / the generated field; ftz)
set -u
mkdir -p gpurun_out
{
for flags in "" "-DCTX_NORM_SKIP_CONST=0" "-DCTX_FAST_DIV=1" "-DCTX_FAST_DIV=2" "-DCTX_FAST_DIV=3" "-DCTX_FAST_DIV=3 --ftz=true" "--ftz=true"; do
echo "== NVRTC flags: [$flags]"; CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench.py --kernel 0 --dtype float32 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k: d[k] for k in ('T', 'ms', 'steps', 'tail')})
"
done
for flags in "" "-DCTX_FAST_DIV=3" "-DCTX_FAST_DIV=3 --ftz=true"; do
echo "== K2 f32, NVRTC flags: [$flags]"; CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench_nuts.py f32 2>&1 | tail -2 | cut -c1-300
done
} > gpurun_out/r2s_div.log 2>&1
cat gpurun_out/r2s_div.log
