#!/bin/bash
# Under gpurun: plain runs first, then ncu (launch list of the bench command + one full capture
# per dominant kernel).  Everything lands in gpurun_out/.
set -u
TAG=${1:-r1}
ARGS="--steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-nuts --no-other"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu_launches.log 2>&1
python bench.py $ARGS > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctx_solve_v[12] -s 12 -c 1 \
    -o gpurun_out/${TAG}_solve_full python bench.py $ARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
python scripts/dev_bench_nuts.py f32 > gpurun_out/${TAG}_nuts_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctx_loglik_v2 -s 2 -c 1 \
    -o gpurun_out/${TAG}_loglik_full python scripts/dev_bench_nuts.py f32 > gpurun_out/${TAG}_nuts_ncu.log 2>&1
python scripts/dev_bench5.py 262144 3 > gpurun_out/${TAG}_mlp_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctx_solve_mlp_tc -s 1 -c 1 \
    -o gpurun_out/${TAG}_mlp_tc_full python scripts/dev_bench5.py 262144 3 > gpurun_out/${TAG}_mlp_ncu.log 2>&1
python scripts/dev_bench3_stab.py > gpurun_out/${TAG}_cfg3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctx_solve_w -s 3 -c 1 \
    -o gpurun_out/${TAG}_solve_w_full python scripts/dev_bench3_stab.py > gpurun_out/${TAG}_cfg3_ncu.log 2>&1
python scripts/dev_bench.py --T 16 --kernel 0 --dtype float32 > gpurun_out/${TAG}_t16_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctx_solve_v1 -s 2 -c 1 \
    -o gpurun_out/${TAG}_v1_T16 python scripts/dev_bench.py --T 16 --kernel 0 --dtype float32 > gpurun_out/${TAG}_t16_ncu.log 2>&1
tail -n 1 gpurun_out/${TAG}_plain.log | cut -c1-300; tail -n 1 gpurun_out/${TAG}_nuts_plain.log | cut -c1-200; tail -n 2 gpurun_out/${TAG}_mlp_plain.log | cut -c1-200
