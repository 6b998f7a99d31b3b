#!/bin/bash
# every configuration of SURVEY.md section 8d through the C ABI at N = 2^28: quick timings (best / median of 5)
mkdir -p gpurun_out
for p in dot dotf32 sum sumi32 multi minloc minloc64 triad triadf32 daxpy daxpyf32 scan exscan scan64 inplacescan64 scanf64 exscanf64 maxscanf32 sort sort_lowent sortdesc sortkeys sortf64 sorti64pairs hist256 hist64k histzipf pi; do
  timeout 300 python tools/run_path.py --path $p --reps 5 2>&1 | tail -1
done > gpurun_out/paths_c.txt 2>&1
cat gpurun_out/paths_c.txt
