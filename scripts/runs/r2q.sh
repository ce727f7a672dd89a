#!/bin/bash
# round 2, GPU call Q (1 GPU): is the critical path of the tail arithmetic latency?  fast-math A/B (flags only),
# new golden GPU tests, bench with the pipelined legs
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_posterior_callers_golden.py tests/test_gpu_posterior_callers.py -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2q_pytest.log
{
for flags in "" "--use_fast_math" "--prec-sqrt=false" "--ftz=true" "--prec-div=false --prec-sqrt=false --ftz=true"; do
echo "== NVRTC flags: [$flags]"; CTX_B200_NVRTC_FLAGS="$flags" python scripts/dev_bench.py --kernel 0 --dtype float32 2>&1 | tail -2 | cut -c1-200
done
} > gpurun_out/r2q_fastmath.log 2>&1
cat gpurun_out/r2q_fastmath.log
python bench.py > gpurun_out/r2q_bench.log 2>gpurun_out/r2q_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2q_bench.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2q_bench.log").read().strip().splitlines()[-1])
for k in ("ms_per_step", "cost_ordered", "pipelined", "pipelined_cost_ordered"):
    v = l.get(k)
    print(k, v if not isinstance(v, dict) else {a: v[a] for a in ("ms_per_step", "hbm_frac")})
print("frac", l["roofline"]["frac"], "launches", l["gpu_launches"])
PY
