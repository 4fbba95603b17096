#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: opcode mix and hot SASS regions."""
import csv, sys
from collections import defaultdict

def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    ia, ie, isamp, it = (hdr.index(k) for k in ("Source", "Instructions Executed", "# Samples", "Avg. Threads Executed"))
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((r[ia].strip(), int(r[ie]), int(r[isamp]), float(r[it])))
        except (ValueError, IndexError):
            continue
    tot = sum(d[1] for d in data); tots = max(1, sum(d[2] for d in data))
    print("total warp-inst", tot, "samples", tots, "sass lines", len(data))
    g = defaultdict(lambda: [0, 0])
    for s, e, sm, t in data:
        parts = s.split()
        op = parts[1] if parts[0].startswith('@') else parts[0]
        g[op.split('.')[0]][0] += e; g[op.split('.')[0]][1] += sm
    for op, (e, sm) in sorted(g.items(), key=lambda x: -x[1][0])[:22]:
        print(f"  {op:10s} inst {e/tot*100:5.1f}%  samples {sm/tots*100:5.1f}%")
    i = 0
    while i < len(data):
        j = i
        while j + 1 < len(data) and abs(data[j + 1][1] - data[i][1]) <= 0.03 * max(data[i][1], 1): j += 1
        n = j - i + 1
        e = sum(d[1] for d in data[i:j + 1]); sm = sum(d[2] for d in data[i:j + 1])
        if e / tot > 0.01:
            print(f"  sass[{i}:{j+1}] n={n} exec/inst={data[i][1]} share={e/tot*100:.1f}% samples={sm/tots*100:.1f}% thr={sum(d[3] for d in data[i:j+1])/n:.1f}")
        i = j + 1

if __name__ == "__main__":
    main(sys.argv[1])
