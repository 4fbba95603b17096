#!/usr/bin/env python
"""Per-kernel launch count / average duration / share of the step from an
`ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_shares.py launches.csv"""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(list)
for r in rows[hi + 1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except (ValueError, IndexError):
        continue
    unit = r[mu]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[r[kn]].append(us)
agg = {k: v for k, v in agg.items() if "fma_probe" not in k}
tot = sum(sum(v) for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 3 --warmup 3 --no-cpu-baseline")
print("# (cold-cache, serialised launches: compare SHARES, not absolutes; the FMA probe kernel is excluded from the shares)")
print("# kernel | launches | avg us | share")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:50]:50s} {len(v):6d} {sum(v)/len(v):9.1f} {sum(v)/tot*100:6.1f}%")
