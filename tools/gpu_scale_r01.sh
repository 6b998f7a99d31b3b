#!/bin/bash
# N-GPU bench line (N = first argument), the way the driver launches it
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r01b_n$N.json 2> gpurun_out/bench_r01b_n$N.err
tail -c 3000 gpurun_out/bench_r01b_n$N.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_r01b_n$N.err | tail -8
