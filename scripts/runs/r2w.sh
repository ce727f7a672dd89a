#!/bin/bash
# round 2, GPU call W (1 GPU): f32 controller without the division / sqrt / reciprocal (same accept decisions)
set -u
mkdir -p gpurun_out
{
for flags in "" "-DCTX_FUSED_FACTOR=0"; do
echo "== f32, NVRTC flags: [$flags]"; CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench.py --kernel 0 --dtype float32 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k: d[k] for k in ('T', 'ms', 'steps', 'tail')})
"
echo "== K2 f32 [$flags]"; CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench_nuts.py f32 2>&1 | tail -2 | cut -c1-300
done
} > gpurun_out/r2w_fused.log 2>&1
cat gpurun_out/r2w_fused.log
python -m pytest tests/test_gpu_parity_solve.py tests/test_gpu_reference_pins.py tests/test_gpu_reference_simulation.py tests/test_gpu_edge_cases.py tests/test_gpu_parity_loglik.py tests/test_gpu_facade.py -m gpu -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2w_pytest.log
