#!/usr/bin/env python
"""Where the slot-pool kernels spend their issue slots: joins `ncu --page source --csv` (per-SASS-address
counters of one profiled launch) with `nvdisasm -gi` of the in-tree libuivr.so (must be the build that was
profiled) and aggregates by the source regions of csrc/uivr_pool.cuh (found by their marker comments; inlined
callees are attributed to the pool line that called them).

    python scripts/ncu_pool_regions.py REPORT.ncu-rep KERNEL_SUBSTR [section_index]
    e.g.  ... gpurun_out/prof.ncu-rep "k_poolILi0ELb0" 0
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
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from pool_regions import MARKS, SRC  # noqa: E402

LIB = os.environ.get("UIVR_LIB") or os.path.join(os.path.dirname(SRC), "libuivr.so")


def regions():
    lines = open(SRC).read().splitlines()
    pos = [(name, next(i + 1 for i, l in enumerate(lines) if mark in l)) for name, mark in MARKS]
    return [(pos[i][1], pos[i + 1][1] - 1, pos[i][0]) for i in range(len(pos) - 1)]


def line_map(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    cur_fn, cur = None, None
    out = {}
    for ln in sass.splitlines():
        m = re.match(r"\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            continue
        if "//## File" in ln:
            ms = re.findall(r'uivr_pool.cuh", line (\d+)', ln)
            cur = int(ms[-1]) if ms else None
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur_fn and kernel_substr in cur_fn:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    lm = line_map(ksub)
    regs = regions()
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    allrows = list(csv.reader(io.StringIO(txt)))
    starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
    print("sections:", [allrows[i][1][:60] for i in starts[:-1]], "-> using", skip)
    rows = allrows[starts[skip]:starts[skip + 1]]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {k: hdr.index(k) for k in ("Address", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
    stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = defaultdict(lambda: [0, 0, 0, 0])  # warp-inst, thread-inst, samples, static
    stalls = defaultdict(lambda: defaultdict(int))
    base = None
    tot = [0, 0, 0]
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
        if base is None:
            base = addr
        line = lm.get(addr - base, (None, ""))[0]
        name = "other"
        if line is not None:
            for a, b, n in regs:
                if a <= line <= b:
                    name = n
        v = [int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0),
             int(r[ci["# Samples"]] or 0)]
        for i in range(3):
            agg[name][i] += v[i]
            tot[i] += v[i]
        agg[name][3] += 1
        for i, sname in stall_cols:
            stalls[name][sname] += int(r[i] or 0)
    print(f"total warp-inst {tot[0]:.3e}  lanes/inst {tot[1] / max(tot[0], 1):.1f}  samples {tot[2]}")
    print(f"{'region':14s} {'warp-inst%':>10s} {'lanes':>6s} {'samples%':>9s} {'static':>7s}  top stalls")
    for name, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        st = sorted(stalls[name].items(), key=lambda kv: -kv[1])[:3]
        ssum = sum(stalls[name].values()) or 1
        print(f"{name:14s} {100 * v[0] / tot[0]:10.2f} {v[1] / max(v[0], 1):6.1f} {100 * v[2] / max(tot[2], 1):9.2f} {v[3]:7d}  "
              + ", ".join(f"{k} {100 * c / ssum:.0f}%" for k, c in st))


if __name__ == "__main__":
    main()
