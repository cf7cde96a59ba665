#!/usr/bin/env python3
"""usage: ncu_summary.py <report.ncu-rep | raw-page.csv>: key metrics + warp-stall breakdown per profiled launch (reads the raw page)."""
import csv, subprocess, sys
out = open(sys.argv[1]).read() if sys.argv[1].endswith('.csv') else \
    subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg.per_second']
for r in rows[2:]:
    for k in keys:
        if k in h:
            print(f"{k} = {r[h.index(k)]} {rows[1][h.index(k)]}")
    st = []
    for i, c in enumerate(h):
        if c.startswith('smsp__average_warps_issue_stalled_') and c.endswith('_per_issue_active.ratio') or \
           c.startswith('smsp__average_warp_latency_issue_stalled_') and c.endswith('.ratio'):
            try:
                st.append((float(r[i]), c.split('stalled_')[1].split('_per_')[0].replace('.ratio', '')))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    for v, n in sorted(st, reverse=True)[:10]:
        print(f"   stall {n:28s} {100 * v / tot:5.1f} %")
    print('----')
