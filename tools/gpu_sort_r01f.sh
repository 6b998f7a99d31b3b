#!/bin/bash
# sort tuning pass f: prefetch distance, one vs two look-back warps, ITEMS 32 at 3 CTAs/SM
mkdir -p gpurun_out
timeout 600 build/tune/tune_sort 28 > gpurun_out/tune_sort_g.txt 2>&1
cat gpurun_out/tune_sort_g.txt
timeout 300 build/tune/tune_sort 28 0xff > gpurun_out/tune_sort_g_lowent.txt 2>&1
cat gpurun_out/tune_sort_g_lowent.txt
