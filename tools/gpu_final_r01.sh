#!/bin/bash
# round-1 final measurement pass (N=1): gpu tests, smoke, bench line, reference arm, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_final.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; tail -3 gpurun_out/bench_r01f.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01f_reference.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_f.log 2>&1
python - <<'P'
import json
for f in ('gpurun_out/bench_r01f.json','gpurun_out/bench_r01f_reference.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(d.get('impl','b200'), d['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'), d.get('paths',{}).get('sort_pairs_u32_u32'), d.get('e2e',{}).get('value'), d.get('gpu_launches'))
P
for p in dot sum multi triad daxpy scan scan64 scanf64 sort sort_lowent sortf64 hist256 hist64k; do
  timeout 300 python tools/run_path.py --path $p --reps 5 2>&1 | tail -1
done > gpurun_out/paths_final.txt 2>&1
cat gpurun_out/paths_final.txt
