#!/bin/bash
# round 2, GPU call Y (1 GPU): bit-identity of the zero-RHS skips on edge inputs
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_edge_cases.py -m gpu -q -k "zero_rhs" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2y_pytest.log
