#!/bin/bash
# round 2, GPU call E (2 GPUs): remaining test files on one GPU, then the 2-rank bench + PCIe ceiling
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_xla_shim.py tests/test_gpu_reference_simulation.py tests/test_gpu_neural_training.py tests/test_gpu_posterior_callers.py -m gpu -q -s > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|mae before" gpurun_out/r2e_pytest.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2e_bench_n2.log 2>&1; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r2e_bench_n2.log
PCIE_GB=12.6 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/pcie_ceiling.py > gpurun_out/r2e_pcie_n2.log 2>&1
tail -2 gpurun_out/r2e_pcie_n2.log
