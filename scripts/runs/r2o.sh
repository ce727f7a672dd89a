#!/bin/bash
# round 2, GPU call O (1 GPU): the scheduling probe -- bit identity, then with / without timings
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sched_probe.py tests/test_gpu_edge_cases.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2o_pytest.log
{
echo "== probe on";  python scripts/dev_bench.py --kernel 0 --dtype float32 float64 2>&1 | tail -4
echo "== probe off"; python scripts/dev_bench.py --kernel 0 --dtype float32 float64 --no-probe 2>&1 | tail -4
for cfg in "12 1e-1" "16 1e-2" "16 1e-1" "32 1e-2" "48 1e-2" "24 1e-3"; do set -- $cfg
echo "== probe cap=$1 rtol=$2"; CTX_B200_PROBE_CAP=$1 CTX_B200_PROBE_RTOL=$2 python scripts/dev_bench.py --kernel 0 --dtype float32 2>&1 | tail -2
done
echo "== uniform costs (K_m 20..50), probe on";  python scripts/dev_bench.py --kernel 0 --dtype float32 --km 20 50 2>&1 | tail -2
echo "== uniform costs, probe off"; python scripts/dev_bench.py --kernel 0 --dtype float32 --km 20 50 --no-probe 2>&1 | tail -2
echo "== hint (upper bound)"; python scripts/dev_bench.py --kernel 0 --dtype float32 --hint 2>&1 | tail -2
} > gpurun_out/r2o_probe.log 2>&1
cat gpurun_out/r2o_probe.log
