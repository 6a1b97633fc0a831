#!/usr/bin/env python
"""Static SASS census of the slot-pool kernels by source region (no GPU needed).  The regions are found by
their marker comments in csrc/uivr_pool.cuh, so the table survives edits.
    python scripts/pool_regions.py [lib.so] [kernel-substring ...]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "unbiased-inverse-volume-rendering_b200", "csrc", "uivr_pool.cuh")
MARKS = [("w-top", "WALKER WARPS:"), ("pickup", "1. idle lanes pick up"), ("w-idle", "const unsigned m_walk ="),
         ("DDA-loop", "for (int it = 1;; ++it)"), ("flush", "3. hand finished lanes on"),
         ("h-sched+pop", "HANDLER WARPS:"), ("TAP", "if (work == Q_TAP)"),
         ("VERTEX", "} else if ((!HAS_ADJ && work == Q_VERTEX)"),
         ("NEE_END", "} else if (work == Q_NEE_END)"), ("PATH_END", "} else if (work == Q_PATH_END)"),
         ("FETCH", "Q_FREE: next work item"), ("SCATTER-stage", "deferred gradients of one described vertex of a finished path"),
         ("NEE-log", "NEE adjoint from the collision log"), ("SPAWN", "one code site: for the batch of"),
         ("scatter", "gradient scatter of the batch"), ("walk-setup", "set-up of a new free-flight walk"),
         ("end", "#undef PU")]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith((".so", ".cubin")) else \
        os.path.join(os.path.dirname(SRC), "libuivr.so")
    kernels = [a for a in sys.argv[1:] if not a.endswith((".so", ".cubin"))] or \
        ["k_pool<0, false", "k_pool<2, false", "k_pool<3, false"]
    lines = open(SRC).read().splitlines()
    pos = []
    for name, mark in MARKS:
        ln = next(i + 1 for i, l in enumerate(lines) if mark in l)
        pos.append((name, ln))
    spec = ",".join(f"{a}-{pos[i + 1][1] - 1}:{n}" for i, (n, a) in enumerate(pos[:-1]))
    for k in kernels:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "sass_lines.py"), lib, k,
                               "regions=" + spec, "exclude=true>"])


if __name__ == "__main__":
    main()
