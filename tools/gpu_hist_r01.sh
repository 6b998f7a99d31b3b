#!/bin/bash
# ncu full capture of the upfront histogram kernel of the sort (u32 keys)
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:histogram_kernel -s 1 -c 1 -o /tmp/ncu/hist -f python tools/run_path.py --path sort --reps 1 > gpurun_out/ncu_hist.log 2>&1
ncu -i /tmp/ncu/hist.ncu-rep --page raw --csv > gpurun_out/hist_raw.csv 2>/dev/null
ncu -i /tmp/ncu/hist.ncu-rep --page source --csv > gpurun_out/hist_source.csv 2>/dev/null
python tools/ncu_summary.py hist=gpurun_out/hist_raw.csv
python tools/ncu_stalls.py gpurun_out/hist_source.csv 8 | head -30
