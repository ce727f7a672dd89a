#!/bin/bash
# round 2, GPU call AC (1 GPU): bench with the imperfect-hint leg
set -u
mkdir -p gpurun_out
python bench.py --no-e2e --no-nuts --no-cpu-baseline > gpurun_out/r2ac_bench.log 2>gpurun_out/r2ac_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2ac_bench.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2ac_bench.log").read().strip().splitlines()[-1])
print(l["ms_per_step"], {k: v for k, v in l["cost_ordered"].items() if k != "schedule"}, l["pipelined"]["ms_per_step"])
PY
