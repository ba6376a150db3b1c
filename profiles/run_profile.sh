#!/bin/bash
# Run under gpurun (1 GPU).  Produces in gpurun_out/:
#   plain.log            the same command without ncu (must exit 0 first)
#   launches.csv         every kernel launch with its device time (cold-cache, serialised: compare shares)
set -e
CMD="python bench.py --pages 600 --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1 || echo "launch list failed"
tail -c 600 gpurun_out/plain.log
