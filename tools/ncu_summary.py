#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics as JSON + top stall lines from the source page.
usage: ncu_summary.py report.ncu-rep [label]"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else rep
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sass__inst_executed_local_loads", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
out = {}
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        out[k] = f"{vals[i]} {units[i]}".strip()
print(json.dumps({label: out}, indent=1))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: 0 for s in stalls}
recs = []
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    st = {}
    for s in stalls:
        try:
            v = int(r[ix[s]])
        except ValueError:
            v = 0
        tot[s] += v
        if v:
            st[s] = v
    recs.append((n, r[ix["Source"]], st))
total = sum(n for n, _, _ in recs)
print("samples", total, sorted(((k, v) for k, v in tot.items() if v), key=lambda x: -x[1])[:10], file=sys.stderr)
recs.sort(key=lambda x: -x[0])
for n, s_, st in recs[:25]:
    print(n, s_[:70], sorted(st.items(), key=lambda x: -x[1])[:3], file=sys.stderr)
