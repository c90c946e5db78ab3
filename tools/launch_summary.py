"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki])[:90]
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"total {T:.1f} us over {sum(cnt.values())} launches")
for n, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[n]:4d}  avg={v / cnt[n]:8.2f} us  {n}")
