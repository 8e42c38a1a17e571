#!/usr/bin/env python
"""Markdown table of the metrics profiles/README.md quotes, from `ncu -i X.ncu-rep --page raw --csv`.

    python tools/ncu_table.py profiles/r01_render_pass_raw.csv
"""
import csv
import re
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
print("| metric | unit | per captured launch |\n|---|---|---|")
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print("| `%s` | %s | %s |" % (w, units[i], ", ".join(d[i] for d in data)))
stalls = {}
for i, h in enumerate(hdr):
    m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", h)
    if m and not m.group(1).endswith("not_issued"):
        stalls[m.group(1)] = float(data[-1][i].replace(",", "") or 0)
total = sum(stalls.values()) or 1.0
print("| stall samples (last launch) | share | %s |" % ", ".join("%s %.0f%%" % (k, 100 * v / total) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]))
