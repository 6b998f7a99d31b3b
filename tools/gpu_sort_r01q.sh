#!/bin/bash
# sort tuning pass q: one look-back warp (288 threads, 72 registers) with the two-level look-back
mkdir -p gpurun_out
timeout 300 build/tune/tune_sort_tk0 28 > gpurun_out/tune_sort_q.txt 2>&1
timeout 300 build/tune/tune_sort_tk0 28 0xff > gpurun_out/tune_sort_q_lowent.txt 2>&1
cat gpurun_out/tune_sort_q.txt
