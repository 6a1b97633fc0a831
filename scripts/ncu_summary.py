#!/usr/bin/env python
"""Summary of an ncu report: key raw metrics + stall reasons per kernel launch."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared_atom.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in keys:
        if k in d:
            print(f"{k:70s} {d[k]:>20s} {units[hdr.index(k)]}")
    print("---")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
seen = set()
for si in range(len(starts) - 1):
    sec = rows[starts[si]:starts[si + 1]]
    name = sec[0][1]
    h = sec[1]
    if 'stall_wait' not in h or (name, len(sec)) in seen:
        continue
    seen.add((name, len(sec)))
    stall = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    tot = {h[i]: 0 for i in stall}
    for r in sec[2:]:
        if len(r) < len(h):
            continue
        for i in stall:
            tot[h[i]] += int(r[i] or 0)
    s = sum(tot.values()) or 1
    print(name)
    print("  stalls: " + ", ".join(f"{k[6:]} {100 * v / s:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
