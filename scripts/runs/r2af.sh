#!/bin/bash
# round 2, GPU call AF (1 GPU): smoke + neural parity after the adjoint change
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2af_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2af_smoke.log | cut -c1-400
python -m pytest tests/test_gpu_parity_neural.py tests/test_gpu_neural_training.py -m gpu -q > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2af_pytest.log
