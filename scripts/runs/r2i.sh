#!/bin/bash
# round 2, GPU call I (1 GPU): the driver's round-end sequence (smoke, full GPU suite, bench, reference
# arm), then the ncu launch list + full captures, then the tail test
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2i_smoke.log | cut -c1-400
python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2i_pytest.log
python bench.py > gpurun_out/r2i_bench.log 2>&1; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_ref.log 2>&1; echo "ref rc=$?"
bash scripts/profile_all.sh r2i
python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 3 > gpurun_out/r2i_T16_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctx_solve_v1 -s 2 -c 1 \
    -o gpurun_out/r2i_v1_T16 python scripts/dev_bench.py --T 16 --dtype float32 --kernel 2 --iters 3 > gpurun_out/r2i_T16_ncu.log 2>&1
echo "== tail test: ms per 2^20 rows at B = 2^20, 2^21 (T=1000 kernel 5; T=16 kernel 2)" > gpurun_out/r2i_tail.log
for B in 1048576 2097152; do
  python scripts/dev_bench.py --B $B --T 1000 --dtype float32 --kernel 5 --iters 3 2>&1 | head -1 | cut -c1-200 >> gpurun_out/r2i_tail.log
  python scripts/dev_bench.py --B $B --T 16 --dtype float32 --kernel 2 --iters 3 2>&1 | head -1 | cut -c1-200 >> gpurun_out/r2i_tail.log
done
cat gpurun_out/r2i_tail.log
