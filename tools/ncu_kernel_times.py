"""Sums gpu__time_duration per kernel name over the LAST `launches` launches of an ncu --csv launch list."""
import csv
import sys
from collections import defaultdict

path, last = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"].split("(")[0], v))
if last:
    rows = rows[-last:]
acc, cnt = defaultdict(float), defaultdict(int)
for k, v in rows:
    acc[k] += v
    cnt[k] += 1
for k in sorted(acc, key=lambda x: -acc[x]):
    print(f"{acc[k]:10.1f} us  x{cnt[k]:<3d} {k[:100]}")
print(f"{sum(acc.values()):10.1f} us  total over {len(rows)} launches")
