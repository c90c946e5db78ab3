"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time per kernel name."""
import csv
import collections
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rd:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("nsecond", "ns") else (v if r[ui] in ("usecond", "us") else v * 1e3)
    name = r[ki]
    a = agg.setdefault(name, [0.0, 0])
    a[0] += us
    a[1] += 1
tot = sum(a[0] for a in agg.values())
n = sum(a[1] for a in agg.values())
print(f"total {tot:.1f} us over {n} launches")
for name, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:10.1f} us {100 * us / tot:5.1f}%  n={c:4d}  avg={us / c:8.2f} us  {name[:100]}")
