#!/bin/bash
# round 2, GPU call J (1 GPU): cost-ordered scheduling hint -- does it pay?
set -u
mkdir -p gpurun_out
{
for T in 1000 250 16; do
  for k in 2 5; do
    if [ $T = 16 ] && [ $k = 5 ]; then continue; fi
    echo "== T=$T kernel=$k plain"; python scripts/dev_bench.py --T $T --dtype float32 --kernel $k --iters 4 2>&1 | head -1 | cut -c1-260
    echo "== T=$T kernel=$k hinted"; python scripts/dev_bench.py --T $T --dtype float32 --kernel $k --iters 4 --hint 2>&1 | head -1 | cut -c1-260
  done
done
echo "== f64 T=1000 kernel 5"; python scripts/dev_bench.py --T 1000 --dtype float64 --kernel 5 --iters 3 2>&1 | head -1 | cut -c1-260
python scripts/dev_bench.py --T 1000 --dtype float64 --kernel 5 --iters 3 --hint 2>&1 | head -1 | cut -c1-260
echo "== K2"; python scripts/dev_bench_nuts.py f32 2>&1 | tail -2
python scripts/dev_bench_nuts.py f64 2>&1 | tail -2
} > gpurun_out/r2j_hint.log 2>&1
cat gpurun_out/r2j_hint.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2j_smoke.log | cut -c1-300
python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity_loglik.py tests/test_gpu_p2p.py -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
