#!/bin/bash
# round-1 measurement pass c (u16-counter sweep, SERIAL fp64 scan, launch reductions): gpu tests, smoke, bench line,
# reference arm, ncu launch list of the bench command, ncu full captures of the sweep and the fp64 scan kernel
mkdir -p gpurun_out /tmp/ncu
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_c.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; tail -c 3500 gpurun_out/bench_r01c.json; tail -3 gpurun_out/bench_r01c.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01c_reference.json 2>&1; tail -c 600 gpurun_out/bench_r01c_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_c.log 2>&1
tail -1 gpurun_out/bench_under_ncu_c.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_late_kernel -s 5 -c 1 -o /tmp/ncu/sort_v4 -f python tools/run_path.py --path sort --reps 1 > gpurun_out/ncu_sort_v4.log 2>&1
ncu -i /tmp/ncu/sort_v4.ncu-rep --page raw --csv > gpurun_out/sort_v4_raw.csv 2>/dev/null
ncu -i /tmp/ncu/sort_v4.ncu-rep --page source --csv > gpurun_out/sort_v4_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o /tmp/ncu/scanf64_v4 -f python tools/run_path.py --path scanf64 --reps 1 > gpurun_out/ncu_scanf64_v4.log 2>&1
ncu -i /tmp/ncu/scanf64_v4.ncu-rep --page raw --csv > gpurun_out/scanf64_v4_raw.csv 2>/dev/null
for p in dot sum multi triad daxpy scan scan64 scanf64 sort sort_lowent sortf64 hist256 hist64k; do
  timeout 300 python tools/run_path.py --path $p --reps 5 2>&1 | tail -1
done > gpurun_out/paths_c2.txt 2>&1
cat gpurun_out/paths_c2.txt
