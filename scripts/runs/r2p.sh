#!/bin/bash
# round 2, GPU call P (1 GPU): new pins on the GPU (log-joint golden), opt-in probe, training with the new penalties
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_logjoint_golden.py tests/test_gpu_sched_probe.py tests/test_gpu_neural_training.py tests/test_gpu_parity_surrogate.py -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2p_pytest.log
