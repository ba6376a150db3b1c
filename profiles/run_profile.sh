#!/bin/bash
# Run under gpurun (1 GPU).  Produces in gpurun_out/ what profiles/README.md cites (copy the summaries into profiles/):
#   r2_plain.log                      the same command without ncu (must exit 0 first)
#   r2_launches_bench_600pages.csv    every kernel launch with its device time (cold-cache, serialised: compare shares)
#   r2_decode_c2_3000.ncu-rep ...     `--set full` captures of the two decode kernels (C2 3,000 pages, C3 400,000 records)
#   r2_lpf.ncu-rep                    `--set full` capture of k_lpf (4th launch of the setitem run)
# Read the reports in the build container:  ncu -i X.ncu-rep --page raw --csv > X_raw.csv
#                                           ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_src.csv
#                                           python profiles/ncu_lines.py X_src.csv <tiles> 70     (per source line)
set -e
CMD="python bench.py --mode setitem --pages 600 --steps 1 --warmup 1 --no-cpu --no-read-side"
$CMD > gpurun_out/r2_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r2_launches_bench_600pages.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1 || echo "launch list failed"
ncu --set full --clock-control none --import-source on -k regex:k_lpf -s 3 -c 1 -f -o gpurun_out/r2_lpf $CMD > gpurun_out/r2_ncu_lpf.log 2>&1 || echo "ncu k_lpf failed"
python profiles/prof_decode.py c2 3000 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_decode_ -s 4 -c 2 -f -o gpurun_out/r2_decode_c2_3000 python profiles/prof_decode.py c2 3000 > gpurun_out/r2_ncu_c2.log 2>&1 || echo "ncu c2 failed"
python profiles/prof_decode.py c3 400000 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_decode_ -s 4 -c 2 -f -o gpurun_out/r2_decode_c3_400k python profiles/prof_decode.py c3 400000 > gpurun_out/r2_ncu_c3.log 2>&1 || echo "ncu c3 failed"
tail -c 600 gpurun_out/r2_plain.log
