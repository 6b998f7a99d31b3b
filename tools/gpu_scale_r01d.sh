#!/bin/bash
# N-GPU bench line (N = first argument), the way the driver launches it; v4 kernels, aligned split sweep
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_r01d_n$N.json 2> gpurun_out/bench_r01d_n$N.err
python - $N <<'P'
import json, sys
for l in open('gpurun_out/bench_r01d_n%s.json' % sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['paths'].items()}, d['kernels']['onesweep_kernel'])
P
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_r01d_n$N.err | tail -5
