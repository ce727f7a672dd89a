#!/bin/bash
# round 2, GPU call U (1 GPU): the driver's sequence on the final code (smoke, full suite, bench, reference arm)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2u_smoke.log | cut -c1-200
python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2u_pytest.log
python bench.py > gpurun_out/r2u_bench.log 2>gpurun_out/r2u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2u_bench.log").read().strip().splitlines()[-1])
for k in ("ms_per_step", "cost_ordered", "pipelined", "pipelined_cost_ordered"):
    v = l.get(k)
    print(k, v if not isinstance(v, dict) else {a: v[a] for a in ("ms_per_step", "hbm_frac")})
print("frac", l["roofline"]["frac"], "launches", l["gpu_launches"], "e2e", l["e2e"]["value"], "nuts", l["nuts"]["value"], l["nuts"].get("cost_ordered", {}).get("value"))
print({k: (v.get("ms"), v.get("value")) for k, v in l["other_configs"].items() if isinstance(v, dict)})
PY
python bench.py --dtype f64 --steps 30 --no-e2e --no-nuts --no-cpu-baseline > gpurun_out/r2u_bench_f64.log 2>&1; echo "f64 rc=$?"
