#!/usr/bin/env python
"""Per-kernel times of the LAST training step in an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/bench_train.py --mode eager."""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi, gi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size')
L = [(r[ki], float(r[vi].replace(',', '')), r[gi]) for r in rows[hdr + 1:] if len(r) > vi]
short = lambda n: re.sub(r'\(.*', '', re.sub(r'void ', '', n))[:60]
last = [i for i, (n, _, _) in enumerate(L) if 'march' in n][-1]
tot, agg = 0.0, OrderedDict()
for n, v, g in L[last:]:
    if '-v' in sys.argv:
        print(f"{v / 1000:8.1f} us  {g:18s} {short(n)}")
    tot += v
    k = short(n)
    agg[k] = (agg.get(k, (0, 0))[0] + v, agg.get(k, (0, 0))[1] + 1)
for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v / 1000:8.1f} us  x{c:<3d} {k}")
print(f"total {tot / 1000:.1f} us over {len(L[last:])} launches")
