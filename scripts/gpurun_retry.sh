#!/bin/bash
# usage: gpurun_retry.sh <logfile> <timeout> [--gpus N] -- <command>   (retries while the pod answers busy)
log=$1; shift; tmo=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $tmo "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry done rc=$rc" >> $log
