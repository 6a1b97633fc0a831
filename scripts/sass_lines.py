#!/usr/bin/env python
"""Static SASS census of one kernel of libuivr.so: instructions per source line of a file, with
inlined callees attributed to the outermost line of that file (nvdisasm -gi).

    python scripts/sass_lines.py <lib.so|cubin> <kernel-substring> [file=uivr_pool.cuh] [regions="a-b:name,..."]

No GPU needed: this is how a change to the walker loop is checked before GPU time is spent on it.
"""
import os
import re
import subprocess
import sys
import tempfile


def cubin_of(path):
    if path.endswith(".cubin"):
        return path
    d = tempfile.mkdtemp(prefix="sass_")
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, stdout=subprocess.DEVNULL)
    return os.path.join(d, sorted(os.listdir(d))[0])


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    opts = dict(a.split("=", 1) for a in sys.argv[3:])
    fname = opts.get("file", "uivr_pool.cuh")
    cubin = cubin_of(lib)
    txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
    # sections
    start = None
    out = {}
    total = 0
    cur = None
    for ln in txt:
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            start = name if (pat in name and (opts.get("exclude") is None or opts["exclude"] not in name)) else None
            if start:
                print("kernel:", name)
            continue
        if start is None:
            continue
        if "//## File" in ln:
            # the LAST "<fname>, line N" on the line is the outermost frame in that file
            ms = re.findall(r'%s", line (\d+)' % re.escape(fname), ln)
            cur = int(ms[-1]) if ms else None
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            total += 1
            out[cur] = out.get(cur, 0) + 1
    print("total SASS instructions:", total, f"({total * 16 / 1024:.1f} KB)")
    if "regions" in opts:
        for part in opts["regions"].split(","):
            rng, nm = part.split(":")
            a, b = map(int, rng.split("-"))
            n = sum(c for l, c in out.items() if l is not None and a <= l <= b)
            print(f"  {nm:16s} lines {a}-{b}: {n}")
    else:
        for l in sorted(k for k in out if k is not None):
            print(l, out[l])
        print("other", out.get(None, 0))


if __name__ == "__main__":
    main()
