#!/bin/bash
# sort tuning pass e: lane-private histogram, packed key+value scatter with late value loads, key prefetch
mkdir -p gpurun_out /tmp/ncu
timeout 600 build/tune/tune_sort 28 > gpurun_out/tune_sort_e.txt 2>&1
cat gpurun_out/tune_sort_e.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:histogram_kernel -c 3 build/tune/tune_sort 28 2>&1 | grep -A3 "histogram_kernel\|gpu__time" | head -20 > gpurun_out/hist_e.txt
cat gpurun_out/hist_e.txt
