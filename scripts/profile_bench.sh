#!/bin/bash
# Run under gpurun: plain bench first, then the ncu launch list and one full capture of the
# dominant kernel (B200_PROFILING.md recipe).  Outputs land in gpurun_out/.
set -u
TAG=${1:-r1}
ARGS="--steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-nuts"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu_launches.log 2>&1
python bench.py $ARGS > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctx_solve_v1 -s 12 -c 1 \
    -o gpurun_out/${TAG}_solve_full python bench.py $ARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -n 2 gpurun_out/${TAG}_plain.log
