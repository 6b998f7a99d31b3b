#!/bin/bash
# round-1 measurement pass b (onesweep v3): gpu tests, smoke, bench line, reference arm, ncu launch
# list of the bench command, ncu full capture of the sweep kernel (text exports only)
mkdir -p gpurun_out /tmp/ncu
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -c 3500 gpurun_out/bench_r01b.json; tail -3 gpurun_out/bench_r01b.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01b_reference.json 2>&1; tail -c 600 gpurun_out/bench_r01b_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_late_kernel -s 5 -c 1 -o /tmp/ncu/sort_v3 -f python tools/run_path.py --path sort --reps 1 > gpurun_out/ncu_sort_v3.log 2>&1
ncu -i /tmp/ncu/sort_v3.ncu-rep --page raw --csv > gpurun_out/sort_v3_raw.csv 2>/dev/null
ncu -i /tmp/ncu/sort_v3.ncu-rep --page source --csv > gpurun_out/sort_v3_source.csv 2>/dev/null
for p in dot sum multi triad daxpy scan scan64 scanf64 sort sort_lowent sortf64 hist256 hist64k; do
  timeout 300 python tools/run_path.py --path $p --reps 5 2>&1 | tail -1
done > gpurun_out/paths_b.txt 2>&1
cat gpurun_out/paths_b.txt
