#!/bin/bash
# forall launch-latency pass d: host-computed tile math, ragged-tile chunks spread over blocks, no second unroll for opaque bodies
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 tests/cpp/_build/bench_launch_latency | tee gpurun_out/bench_launch_latency.txt
for p in dot sum multi triad daxpy hist64k; do timeout 300 python tools/run_path.py --path $p --reps 5 2>&1 | tail -1; done | tee gpurun_out/paths_d.txt
timeout 600 tests/cpp/_build/bench_adapter 28 2>&1 | tee gpurun_out/adapter_bench_d.txt | head -30
timeout 300 tests/cpp/_build/bench_workgroup 2>&1 | tee gpurun_out/bench_workgroup_d.txt
