#!/bin/bash
# sort tuning pass h: sweep shapes for every (key, value) width pair (tools/tune/tune_sort.cu -DTK=0..5)
mkdir -p gpurun_out
for tk in 0 1 2 3 4 5; do
  echo "== TK=$tk"
  timeout 300 build/tune/tune_sort_tk$tk 28 2>&1
done > gpurun_out/tune_sort_h.txt
cat gpurun_out/tune_sort_h.txt
