#!/bin/bash
# round 2, GPU call K (1 GPU): which fraction of the rows should the cost-order hint move to the front?
set -u
mkdir -p gpurun_out
{
for frac in 0.004 0.016 0.0625 0.25 1.0; do
  export CTX_B200_COST_ORDER_FRAC=$frac
  echo "== frac=$frac T=1000 v2"; python scripts/dev_bench.py --T 1000 --dtype float32 --kernel 5 --iters 4 --hint 2>&1 | head -1 | cut -c1-120
  echo "== frac=$frac T=250 v2"; python scripts/dev_bench.py --T 250 --dtype float32 --kernel 5 --iters 4 --hint 2>&1 | head -1 | cut -c1-120
  echo "== frac=$frac T=16 v1"; python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 4 --hint 2>&1 | head -1 | cut -c1-120
  echo "== frac=$frac K2 f32"; python scripts/dev_bench_nuts.py f32 2>&1 | tail -2 | head -1
done
} > gpurun_out/r2k_frac.log 2>&1
cat gpurun_out/r2k_frac.log
