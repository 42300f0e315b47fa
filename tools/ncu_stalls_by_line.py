#!/usr/bin/env python3
"""ncu --page source --csv (one kernel) -> stall samples grouped by source line and by what the line does.
usage: ncu -i rep.ncu-rep --page source --csv -k <kernel> | python tools/ncu_stalls_by_line.py [top_n]
Groups: the warp-stall sampling columns of the source page ("# Samples" per stall reason when present, else the
total sampling column) summed per source file:line; prints the top lines and the share of every file."""
import csv
import collections
import re
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
src_col = next((i for i, h in enumerate(hdr) if h.strip().lower() in ("source", "#", "address")), 0)
samp_cols = [i for i, h in enumerate(hdr) if "Samples" in h or h.startswith("stall_") or "Stall" in h]
line_col = next((i for i, h in enumerate(hdr) if h.strip().lower() in ("source file", "file", "source location", "location")), None)
per_line = collections.Counter()
per_reason = collections.Counter()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    loc = r[line_col] if line_col is not None else r[src_col]
    tot = 0.0
    for i in samp_cols:
        try:
            v = float(r[i].replace(",", "") or 0)
        except ValueError:
            continue
        per_reason[hdr[i]] += v
        if hdr[i].strip() in ("# Samples", "Samples", "Warp Stall Sampling (All Samples)"):
            tot = max(tot, v)
    per_line[loc] += tot
total = sum(per_line.values()) or 1.0
print("# columns:", ", ".join(hdr))
print("# stall samples by source location (share of all samples)")
for loc, v in per_line.most_common(top_n):
    print(f"{100 * v / total:6.2f}%  {loc}")
print("# by reason / column")
for k, v in per_reason.most_common():
    print(f"{v:12.0f}  {k}")
