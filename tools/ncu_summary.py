#!/usr/bin/env python3
"""ncu --page raw --csv  ->  one compact row per profiled launch (the summary committed under profiles/).
usage: ncu -i report.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/ncu_<round>_full_summary.csv"""
import csv
import re
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]

rows = list(csv.reader(line for line in sys.stdin if not line.startswith("==")))
hdr, units = rows[0], rows[1]
want = [w for w in WANT if w in hdr]
idx = [hdr.index(w) for w in want]
out = csv.writer(sys.stdout)
out.writerow(["# ncu --set full --clock-control none; one profiled launch per row"])
out.writerow(["unit"] + [units[i] for i in idx[1:]])
out.writerow(want)
seen = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx[0]]).replace("void ", "")
    key = (name, r[idx[1]])
    if seen.get(key, 0) >= 2:        # two samples per (kernel, grid) are enough
        continue
    seen[key] = seen.get(key, 0) + 1
    out.writerow([name] + [r[i] for i in idx[1:]])
