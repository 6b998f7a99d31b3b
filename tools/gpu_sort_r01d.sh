#!/bin/bash
# sort tuning pass d: late-count kernel with the two-stage look-back; ncu text exports of the best candidate
mkdir -p gpurun_out /tmp/ncu
timeout 600 build/tune/tune_sort 28 > gpurun_out/tune_sort_d.txt 2>&1
cat gpurun_out/tune_sort_d.txt
timeout 300 build/tune/tune_sort 28 0xff > gpurun_out/tune_sort_d_lowent.txt 2>&1
cat gpurun_out/tune_sort_d_lowent.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_late_kernel -s 5 -c 1 -o /tmp/ncu/sort_late -f build/tune/tune_sort 27 > gpurun_out/ncu_sort_late.log 2>&1
ncu -i /tmp/ncu/sort_late.ncu-rep --page raw --csv > gpurun_out/sort_late_raw.csv 2>/dev/null
ncu -i /tmp/ncu/sort_late.ncu-rep --page source --csv > gpurun_out/sort_late_source.csv 2>/dev/null
