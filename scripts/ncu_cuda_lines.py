#!/usr/bin/env python
"""Hottest CUDA source lines of one kernel from `ncu -i REP --page source --csv --print-source cuda,sass` (the
report carries its own source correlation: independent of the in-tree build).

    ncu -i REP --page source --csv --print-source cuda,sass > /tmp/src.csv
    python scripts/ncu_cuda_lines.py /tmp/src.csv "k_pool<(int)2" [top]
"""
import csv
import sys
from collections import defaultdict

path, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
agg = defaultdict(lambda: [0, 0, 0, ""])   # samples, warp-inst, thread-inst, text
cur_file, cur_fn, hdr = None, None, None
seen_fn_file = set()
skip = False
for row in csv.reader(open(path, newline="")):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        cur_fn = row[1]
        key = (cur_fn, cur_file)
        skip = key in seen_fn_file   # every launch of the kernel repeats its blocks: keep the first
        seen_fn_file.add(key)
        continue
    if row[0] == "Line No":
        hdr = row
        ci = {"s": hdr.index("# Samples"), "i": hdr.index("Instructions Executed"), "t": hdr.index("Thread Instructions Executed")}
        continue
    if skip or hdr is None or cur_fn is None or ksub not in cur_fn or not row[0] or len(row) != len(hdr) or not row[0].isdigit():
        continue
    def num(v):
        try:
            return int(v)
        except ValueError:
            return 0
    a = agg[(cur_file, int(row[0]))]
    a[0] += num(row[ci["s"]]); a[1] += num(row[ci["i"]]); a[2] += num(row[ci["t"]]); a[3] = row[1].strip()[:110]
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"{ksub}: {tot_s} samples, {tot_i:.3e} warp instructions")
print(f"{'file:line':28s} {'samples%':>8s} {'inst%':>6s} {'lanes':>5s}  source")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f + ':' + str(l):28s} {100 * a[0] / tot_s:8.2f} {100 * a[1] / tot_i:6.2f} {a[2] / max(a[1], 1):5.1f}  {a[3]}")
