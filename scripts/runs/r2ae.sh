#!/bin/bash
# round 2, GPU call AE (1 GPU): max_time gradient of the adjoint kernels; the whole training suite
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_neural_training.py -m gpu -q -k "weight_gradient or max_time or l1" > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2ae_pytest.log | cut -c1-220
