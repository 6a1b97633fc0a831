#!/usr/bin/env python
"""Aggregate scripts/ncu_lines.py output by source regions: REGIONS = "start-end:name,..." """
import re, sys
txt, file, spec = sys.argv[1], sys.argv[2], sys.argv[3]
regions = []
for part in spec.split(","):
    rng, name = part.split(":")
    a, b = rng.split("-")
    regions.append((int(a), int(b), name))
agg = {}
for ln in open(txt):
    m = re.match(r'(\S+):(-?\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)', ln)
    if not m:
        continue
    f, l, wi, thr, sm = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
    name = 'other:' + f
    if f == file:
        for a, b, n in regions:
            if a <= l <= b:
                name = n
    d = agg.setdefault(name, [0, 0, 0])
    d[0] += wi; d[1] += wi * thr; d[2] += sm
for n, d in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{n:20s} warp-inst {d[0]:6.2f}%  avg-thr {d[1] / max(d[0], 1e-9):5.1f}  samples {d[2]:6.2f}%")
