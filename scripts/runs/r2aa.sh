#!/bin/bash
# round 2, GPU call AA (1 GPU): surrogate-mode LOO reconstruction against the reference functions
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_posterior_callers_golden.py tests/test_gpu_posterior_callers.py -m gpu -q > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2aa_pytest.log
