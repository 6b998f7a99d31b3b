#!/bin/bash
# round-1 profiling pass: quick timings of every path + ncu full captures of scan and onesweep
mkdir -p gpurun_out
for p in dot sum multi triad daxpy scan scan64 scanf64 sort sort_lowent sortf64 hist256 hist64k hist256agg hist64kagg; do
  timeout 300 python tools/run_path.py --path $p --reps 5 2>&1 | tail -1
done > gpurun_out/paths_a.txt 2>&1
cat gpurun_out/paths_a.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/scan_a python tools/run_path.py --path scan --reps 1 > gpurun_out/ncu_scan_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 4 -c 1 -o gpurun_out/sort_a python tools/run_path.py --path sort --reps 1 > gpurun_out/ncu_sort_a.log 2>&1
ls -la gpurun_out
