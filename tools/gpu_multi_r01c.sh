#!/bin/bash
# 2-GPU pass c (v4 sweep): multi-GPU parity tests, sort breakdown, bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_gloo.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/micro/multi_sort_breakdown.py > gpurun_out/multi_sort_breakdown_n2_v4.txt 2>&1; tail -12 gpurun_out/multi_sort_breakdown_n2_v4.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r01c_n2.json 2> gpurun_out/bench_r01c_n2.err
tail -c 2500 gpurun_out/bench_r01c_n2.json; tail -5 gpurun_out/bench_r01c_n2.err
