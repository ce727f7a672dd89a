#!/bin/bash
# round 2, GPU call AD (1 GPU): RateFlow / UniversalODE training with the penalty presets
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_neural_training.py -m gpu -q -k "penalty_presets or strategy" > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2ad_pytest.log | cut -c1-200
