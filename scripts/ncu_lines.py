#!/usr/bin/env python
"""Per-source-line aggregation of an ncu SASS-level source page.

    python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_MANGLED_SUBSTR [launch_skip] [top]

Joins `ncu --page source --csv` (per-SASS-address counters) with `nvdisasm -g` line info of the
in-tree libuivr.so (must be the build that was profiled) and prints, per source line: warp
instructions executed, thread instructions, average active threads, stall samples.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unbiased-inverse-volume-rendering_b200", "csrc", "libuivr.so")


def line_map(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    cur_fn, cur_line, inl = None, None, None
    out = {}
    for ln in sass.splitlines():
        m = re.match(r"\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            m2 = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            if m2 and os.environ.get("NCU_LINES_PARENT"):
                cur_line = (cur_line[0] + ":" + str(cur_line[1]) + " <- " + os.path.basename(m2.group(1)), int(m2.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur_fn and kernel_substr in cur_fn:
            out[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    lm = line_map(ksub)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    allrows = list(csv.reader(io.StringIO(txt)))
    starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
    print("sections:", [allrows[i][1] for i in starts[:-1]], "-> using", skip)
    rows = allrows[starts[skip]:starts[skip + 1]]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    samples_col = os.environ.get("NCU_LINES_STALL", "# Samples")  # e.g. stall_no_inst
    ci = {k: hdr.index(k) for k in ("Address", "Source", samples_col, "Instructions Executed",
                                    "Thread Instructions Executed", "Predicated-On Thread Instructions Executed")}
    agg = defaultdict(lambda: [0, 0, 0, 0])
    base = None
    tot = [0, 0, 0, 0]
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
        if base is None:
            base = addr
        off = addr - base
        key = lm.get(off, ((None, -1), ""))[0]
        v = [int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0),
             int(r[ci["Predicated-On Thread Instructions Executed"]] or 0), int(r[ci[samples_col]] or 0)]
        for i in range(4):
            agg[key][i] += v[i]
            tot[i] += v[i]
    print(f"total warp-inst {tot[0]:.3e} thread-inst {tot[1]:.3e} pred-on {tot[2]:.3e} samples {tot[3]}")
    print(f"{'file:line':28s} {'warp-inst%':>10s} {'avg-thr':>8s} {'samples%':>9s}")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][3])[:top]:
        name = f"{key[0]}:{key[1]}"
        print(f"{name:28s} {100 * v[0] / tot[0]:10.2f} {v[1] / max(v[0], 1):8.1f} {100 * v[3] / max(tot[3], 1):9.2f}")


if __name__ == "__main__":
    main()
