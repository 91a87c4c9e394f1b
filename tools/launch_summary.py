"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) count, mean and total."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi, gi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Unit")
agg = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0][-60:]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    key = (name, r[gi])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) * scale
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'grid':>16s} {'launches':>8s} {'mean us':>10s} {'total us':>10s} {'share':>6s}")
for (name, grid), (c, t) in agg.items():
    print(f"{name:60s} {grid:>16s} {c:8d} {t / c:10.1f} {t:10.1f} {100 * t / tot:5.1f}%")
