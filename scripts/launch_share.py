#!/usr/bin/env python
"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) per kernel.

    python scripts/launch_share.py profiles/r02_launches_bench_steps2.csv "note" > profiles/r02_launch_share.txt
"""
import collections
import csv
import io
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(io.StringIO("".join(rows))))
h = r[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
agg = collections.OrderedDict()
for x in r[1:]:
    ms = float(x[vi].replace(",", "")) * scale.get(x[ui], 1e-6)
    a = agg.setdefault(x[ki][:72], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# per-launch times under ncu are cold-cache and serialised: only the shares are comparable")
print(f"{'kernel':74s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:74s} {a[0]:8d} {a[1]:10.3f} {100 * a[1] / tot:6.1f}%")
