#!/usr/bin/env python
"""Side-by-side summary of `ncu --page raw --csv` exports:  ncu_summary.py tag=file.csv [tag=file.csv ...]"""
import csv, sys
NAMES = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread dram__bytes_read.sum dram__bytes_write.sum
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum l1tex__t_requests_pipe_lsu_mem_global_op_st.sum sm__throughput.avg.pct_of_peak_sustained_elapsed
lts__throughput.avg.pct_of_peak_sustained_elapsed gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed""".split()
cols = []
for a in sys.argv[1:]:
    tag, path = a.rsplit("=", 1)
    rows = list(csv.reader(open(path)))
    cols.append((tag, {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}))
for tag, d in cols:
    print(f"# {tag}: {d.get('Kernel Name', ('', ''))[0][:160]}")
print(f"{'':88s}" + "".join(f"{t:>20s}" for t, _ in cols))
for n in NAMES:
    if any(n in d for _, d in cols):
        unit = next((d[n][1] for _, d in cols if n in d), "")
        print(f"{n:88s}" + "".join(f"{d.get(n, ('-', ''))[0]:>20s}" for _, d in cols) + f" {unit}")
