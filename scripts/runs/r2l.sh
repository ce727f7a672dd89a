#!/bin/bash
# round 2, GPU call L (1 GPU): cost-order hint with the strided queue start
set -u
mkdir -p gpurun_out
{
echo "== plain T=1000 v2"; python scripts/dev_bench.py --T 1000 --dtype float32 --kernel 5 --iters 4 2>&1 | head -1 | cut -c1-120
echo "== plain T=250 v2"; python scripts/dev_bench.py --T 250 --dtype float32 --kernel 5 --iters 4 2>&1 | head -1 | cut -c1-120
echo "== plain T=16 v1"; python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 4 2>&1 | head -1 | cut -c1-120
for frac in 0.0625 0.25 1.0; do
  export CTX_B200_COST_ORDER_FRAC=$frac
  echo "== frac=$frac T=1000 v2"; python scripts/dev_bench.py --T 1000 --dtype float32 --kernel 5 --iters 4 --hint 2>&1 | head -1 | cut -c1-120
  echo "== frac=$frac T=250 v2"; python scripts/dev_bench.py --T 250 --dtype float32 --kernel 5 --iters 4 --hint 2>&1 | head -1 | cut -c1-120
  echo "== frac=$frac T=250 v1"; python scripts/dev_bench.py --T 250 --dtype float32 --kernel 2 --iters 4 --hint 2>&1 | head -1 | cut -c1-120
  echo "== frac=$frac T=16 v1"; python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 4 --hint 2>&1 | head -1 | cut -c1-120
done
export CTX_B200_COST_ORDER_FRAC=1.0
echo "== f64 T=1000 v2 hinted"; python scripts/dev_bench.py --T 1000 --dtype float64 --kernel 5 --iters 3 --hint 2>&1 | head -1 | cut -c1-120
echo "== f64 T=16 v1 hinted"; python scripts/dev_bench.py --T 16 --dtype float64 --kernel 2 --iters 3 --hint 2>&1 | head -1 | cut -c1-120
} > gpurun_out/r2l_hint.log 2>&1
cat gpurun_out/r2l_hint.log
python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity_loglik.py -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest.log
