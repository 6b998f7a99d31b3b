#!/bin/bash
# scan tuning pass e (SERIAL shapes) + full gpu parity + bench with the u16-counter sort
mkdir -p gpurun_out
timeout 600 build/tune/tune_scan 28 > gpurun_out/tune_scan_e.txt 2>&1
grep "f64" gpurun_out/tune_scan_e.txt | head -24
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_p.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01p.json 2> gpurun_out/bench_r01p.err
python - <<'P'
import json
for l in open('gpurun_out/bench_r01p.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['paths']['sort_pairs_u32_u32'], d['kernels']['onesweep_kernel'])
P
