#!/bin/bash
# ncu full captures of the late-count (MATCH 3) and early-count (MATCH 2) sweep kernels; only text exports come back
mkdir -p gpurun_out /tmp/ncu
for v in late early; do
  if [ $v = late ]; then k=onesweep_late_kernel; else k=onesweep_kernel; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -o /tmp/ncu/sort_$v -f build/tune/tune_sort 27 > gpurun_out/ncu_sort_$v.log 2>&1
  ncu -i /tmp/ncu/sort_$v.ncu-rep --page raw --csv > gpurun_out/sort_${v}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/sort_$v.ncu-rep --page source --csv > gpurun_out/sort_${v}_source.csv 2>/dev/null
  ncu -i /tmp/ncu/sort_$v.ncu-rep --page details > gpurun_out/sort_${v}_details.txt 2>/dev/null
done
ls -la gpurun_out
