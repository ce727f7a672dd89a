#!/bin/bash
# round 2, GPU call AG (1 GPU): Model.simulate against the reference's real front end
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_simulate_frontend_golden.py -m gpu -q > gpurun_out/r2ag_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2ag_pytest.log | cut -c1-220
