#!/bin/bash
# round-1 measurement pass: gpu tests, bench line, ncu launch list of the same command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 3000 gpurun_out/bench_r01.json; tail -3 gpurun_out/bench_r01.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2>&1; cat gpurun_out/bench_r01_reference.json | tail -c 1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
