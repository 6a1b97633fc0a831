#!/usr/bin/env python
"""profiles/kernel_constants.json from an ncu report of ONE config-3 step (scripts/profile_step.py variant=3,
`ncu --set full -k regex:k_pool -s 3 -c 3`): warp instructions, lanes per instruction and DRAM bytes of the three
slot-pool launches (forward, adjoint, DRT), stamped with the fingerprint of the kernel sources.
bench.py reads it for `roofline.traffic` / `roofline.issue` and refuses it when the sources have changed.

    python scripts/ncu_constants.py gpurun_out/prof.ncu-rep "r02 gpurun call A"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rep, captured = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "?")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def val(d, k):
        u = units[hdr.index(k)]
        v = float(d[k].replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1.0, "usecond": 1e-3, "second": 1e3,
                 "nsecond": 1e-6}.get(u, 1.0)
        return v * scale

    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        launches.append({
            "kernel": d["Kernel Name"],
            "ms_under_ncu": val(d, "gpu__time_duration.sum"),
            "warp_inst": val(d, "smsp__inst_executed.sum"),
            "lanes_per_inst": val(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "issue_active_pct": val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "dram_bytes": val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum"),
            "l1_hit_pct": val(d, "l1tex__t_sector_hit_rate.pct"),
            "regs": val(d, "launch__registers_per_thread"),
        })
    if len(launches) != 3 or "k_pool<0" not in launches[0]["kernel"].replace("(int)", ""):
        raise SystemExit(f"expected the 3 slot-pool launches of one step (forward first), found "
                         f"{[l['kernel'][:40] for l in launches]}")
    fwd, adj, drt = launches
    wi_b = adj["warp_inst"] + drt["warp_inst"]
    out = {
        "source_sha": bench.kernel_source_sha(),
        "captured": captured,
        "what": "one step of bench.py's config 3 (256^3, 512x512x64 spp) under `ncu --set full --clock-control none`: "
                "launches = forward, adjoint replay (gathers L itself), DRT",
        "fwd_warp_inst_per_step": fwd["warp_inst"],
        "bwd_pipeline_warp_inst_per_step": wi_b,
        "fwd_dram_bytes_per_launch": fwd["dram_bytes"],
        "bwd_pipeline_dram_bytes_per_step": adj["dram_bytes"] + drt["dram_bytes"],
        "lanes_per_instruction": {
            "forward": fwd["lanes_per_inst"],
            "backward_pipeline": (adj["lanes_per_inst"] * adj["warp_inst"] + drt["lanes_per_inst"] * drt["warp_inst"]) / wi_b},
        "launches": launches,
    }
    path = os.path.join(ROOT, "profiles", "kernel_constants.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, "source_sha", out["source_sha"])


if __name__ == "__main__":
    main()
