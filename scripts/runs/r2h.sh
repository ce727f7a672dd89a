#!/bin/bash
# round 2, GPU call H (8 GPUs): one-shot NVLink all-reduce check, then the 8-rank bench
set -u
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/p2p_check.py > gpurun_out/r2h_p2p_n8.log 2>&1; echo "p2p rc=$?"
grep -E "^\{|MISMATCH|timed out" gpurun_out/r2h_p2p_n8.log | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r2h_bench_n8.log 2>&1; echo "bench n8 rc=$?"
grep -E "^\{" gpurun_out/r2h_bench_n8.log | tail -c 1500
grep -E "Error|error" gpurun_out/r2h_bench_n8.log | head -5
