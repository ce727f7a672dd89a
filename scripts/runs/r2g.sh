#!/bin/bash
# round 2, GPU call G (2 GPUs): one-shot NVLink all-reduce check, then the 2-rank bench
set -u
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/p2p_check.py > gpurun_out/r2g_p2p_n2.log 2>&1; echo "p2p rc=$?"
grep -E "^\{|MISMATCH|Error|error|timed out" gpurun_out/r2g_p2p_n2.log | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2g_bench_n2.log 2>&1; echo "bench n2 rc=$?"
grep -E "^\{" gpurun_out/r2g_bench_n2.log | tail -c 2500
grep -E "Error|error" gpurun_out/r2g_bench_n2.log | head -5
