#!/bin/bash
# ncu full capture of the fp64 scan kernel (text exports only) next to the int64 one
mkdir -p gpurun_out /tmp/ncu
for p in scanf64 scan64; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o /tmp/ncu/$p -f python tools/run_path.py --path $p --reps 1 > gpurun_out/ncu_$p.log 2>&1
  ncu -i /tmp/ncu/$p.ncu-rep --page raw --csv > gpurun_out/${p}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$p.ncu-rep --page source --csv > gpurun_out/${p}_source.csv 2>/dev/null
done
ls -la gpurun_out | tail -5
