#!/bin/bash
# sort tuning pass b: late-count onesweep variants vs the early-count kernel + ncu full captures of both
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/sort_b_smi.txt
timeout 600 build/tune/tune_sort 28 > gpurun_out/tune_sort_b.txt 2>&1
cat gpurun_out/tune_sort_b.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_late_kernel -s 4 -c 1 -o gpurun_out/sort_late3 -f build/tune/tune_sort 26 > gpurun_out/ncu_sort_late3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 4 -c 1 -o gpurun_out/sort_early2 -f build/tune/tune_sort 26 > gpurun_out/ncu_sort_early2.log 2>&1
tail -3 gpurun_out/ncu_sort_late3.log
ls -la gpurun_out
