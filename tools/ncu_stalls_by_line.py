#!/usr/bin/env python3
"""ncu --page source --print-source cuda,sass --csv  ->  warp-stall samples per CUDA source line and stall reason.
usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv | python tools/ncu_stalls_by_line.py [kernel-substring] [top_n]
The page is a sequence of sections (File Path / Function Name / header / rows); rows with a line number are the
per-source-line aggregates of the SASS rows below them.  A sample is attributed to the instruction the warp could not
issue, so a load's latency shows up on the first instruction that CONSUMES it, with the reason (long_sb = global / local
memory, short_sb / mio = shared memory, barrier, math = pipe busy, wait = fixed latency, not_selected = another warp
issued)."""
import collections
import csv
import os
import sys

want = sys.argv[1] if len(sys.argv) > 1 else ""
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
REASONS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_lg",
           "stall_not_selected", "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst", "stall_misc"]
per_line = collections.defaultdict(lambda: collections.Counter())
src_text = {}
fpath = func = None
hdr = None
seen_funcs = []
for r in csv.reader(l for l in sys.stdin if not l.startswith("==")):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
        continue
    if r[0] == "Function Name":
        func = r[1]
        if func not in seen_funcs:
            seen_funcs.append(func)
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0] or not r[0].isdigit() or want not in (func or ""):
        continue
    if func != next((f for f in seen_funcs if want in f), None):
        continue                                # the first matching launch only (the others repeat it)
    key = (os.path.basename(fpath), int(r[0]))
    src_text[key] = r[1].strip()
    for i, h in enumerate(hdr):
        if i < len(r) and (h in REASONS or h == "# Samples"):
            try:
                per_line[key][h] += float(r[i])
            except ValueError:
                pass
tot = sum(c["# Samples"] for c in per_line.values()) or 1.0
print(f"# kernel: {next((f for f in seen_funcs if want in f), None)}")
print(f"# {int(tot)} warp-stall samples; share per source line, then its samples by reason")
by_reason = collections.Counter()
for c in per_line.values():
    for k in REASONS:
        by_reason[k] += c[k]
print("# all lines, by reason: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in by_reason.most_common() if v))
by_file = collections.Counter()
for (f, _), c in per_line.items():
    by_file[f] += c["# Samples"]
print("# by file: " + ", ".join(f"{f} {100 * v / tot:.1f}%" for f, v in by_file.most_common()))
for key, c in sorted(per_line.items(), key=lambda kv: -kv[1]["# Samples"])[:top_n]:
    rs = ", ".join(f"{k[6:]} {int(c[k])}" for k in REASONS if c[k] >= 0.02 * c["# Samples"] and c[k] > 0)
    print(f"{100 * c['# Samples'] / tot:6.2f}%  {key[0]}:{key[1]:<4d} {src_text[key][:70]:70s} | {rs}")
