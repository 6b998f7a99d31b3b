#!/bin/bash
# 2-GPU pass b: NCCL parity of the sharded paths (partition-first sample sort), then the bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_gloo.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r01b_n2.json 2> gpurun_out/bench_r01b_n2.err
tail -c 2500 gpurun_out/bench_r01b_n2.json; tail -5 gpurun_out/bench_r01b_n2.err
