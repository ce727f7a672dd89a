#!/bin/bash
# round 2, GPU call N (2 GPUs): final state -- smoke, full GPU suite (incl. the 2-GPU P2P test), bench at N=1 and N=2, f64 bench
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2n_smoke.log | cut -c1-200
python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2n_pytest.log
python bench.py > gpurun_out/r2n_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2n_bench_n2.log 2>&1; echo "bench n2 rc=$?"
python bench.py --dtype f64 --steps 30 --no-e2e > gpurun_out/r2n_bench_f64.log 2>&1; echo "bench f64 rc=$?"
