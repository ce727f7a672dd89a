#!/bin/bash
# round 2, GPU call X (1 GPU): full GPU suite on the fused f32 controller + stage skip
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2x_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2x_smoke.log | cut -c1-200
