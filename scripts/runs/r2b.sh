#!/bin/bash
# round 2, GPU call B: smoke, full GPU test suite, full bench, reference arm, T=16 profile of v1
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r2b_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest.log
python bench.py > gpurun_out/r2b_bench.log 2>&1; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2b_bench.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.log 2>&1
python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 3 > gpurun_out/r2b_T16_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctx_solve_v1 -s 2 -c 1 \
    -o gpurun_out/r2b_v1_T16 python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 3 > gpurun_out/r2b_T16_ncu.log 2>&1
echo done
