#!/bin/bash
# Run under gpurun (1 GPU).  Produces in gpurun_out/:
#   plain.log            the same command without ncu (must exit 0 first)
#   launches.csv         every kernel launch with its device time (cold-cache, serialised: compare shares)
#   onesweep.ncu-rep     one `--set full` capture of the dominant kernel (k_rs_onesweep)
set -e
CMD="python bench.py --pages 600 --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1 || echo "launch list failed"
$CMD > gpurun_out/plain2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rs_onesweep -s 40 -c 3 -o gpurun_out/onesweep $CMD > gpurun_out/ncu_full.log 2>&1 || echo "full capture failed"
tail -2 gpurun_out/plain.log
ls -la gpurun_out
