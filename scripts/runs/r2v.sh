#!/bin/bash
# round 2, GPU call V (2 GPUs): bench at N=2 with the pipelined legs, 2-GPU tests
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2v_bench_n2.log 2>gpurun_out/r2v_bench_n2.err; echo "bench n2 rc=$?"
tail -c 400 gpurun_out/r2v_bench_n2.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2v_bench_n2.log").read().strip().splitlines()[-1])
for k in ("ms_per_step", "cost_ordered", "pipelined", "pipelined_cost_ordered"):
    v = l.get(k)
    print(k, v if not isinstance(v, dict) else {a: v[a] for a in ("ms_per_step", "hbm_frac")})
print("n", l["n_gpus"], "value", l["value"], "e2e", l["e2e"]["value"], "nuts", l["nuts"]["value"], l["nuts"].get("series_sharded"))
PY
timeout 600 python -m pytest tests/test_gpu_p2p.py -m gpu -q 2>&1 | tail -3
