#!/bin/bash
# round 2, GPU call AB (8 GPUs): the bench at N=8 on the final code
set -u
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r2ab_bench_n8.log 2>gpurun_out/r2ab_bench_n8.err; echo "bench n8 rc=$?"
tail -c 300 gpurun_out/r2ab_bench_n8.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2ab_bench_n8.log").read().strip().splitlines()[-1])
for k in ("ms_per_step", "cost_ordered", "pipelined", "pipelined_cost_ordered"):
    v = l.get(k)
    print(k, v if not isinstance(v, dict) else {a: v[a] for a in ("ms_per_step", "hbm_frac")})
print("n", l["n_gpus"], "value", l["value"], "e2e", l["e2e"]["value"], l["e2e"]["d2h_ceiling"]["frac_of_ceiling"], "nuts", l["nuts"]["value"])
ss = l["nuts"].get("series_sharded", {})
print({k: ss.get(k) for k in ("value", "ms_per_eval_batch")}, ss.get("breakdown_ms"))
print({k: (v.get("ms"), v.get("value")) for k, v in l["other_configs"].items() if isinstance(v, dict)})
PY
