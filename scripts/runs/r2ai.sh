#!/bin/bash
# round 2, GPU call AI (1 GPU): Model rate function (K4) against the reference's real stacks
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rate_function_golden.py -m gpu -q > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2ai_pytest.log | cut -c1-220
