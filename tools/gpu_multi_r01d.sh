#!/bin/bash
# 2-GPU pass d: line-aligned write-out of the split sweep; parity (1 GPU partition tests, 2-GPU sharded paths), breakdown, bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sort or partition or multi" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/micro/multi_sort_breakdown.py > gpurun_out/multi_sort_breakdown_n2_v5.txt 2>&1; grep " ms" gpurun_out/multi_sort_breakdown_n2_v5.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_r01d_n2.json 2> gpurun_out/bench_r01d_n2.err
python - <<'P'
import json
for l in open('gpurun_out/bench_r01d_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['paths'], d['kernels']['onesweep_kernel'])
P
tail -3 gpurun_out/bench_r01d_n2.err
