#!/usr/bin/env python
"""Per-CUDA-source-line stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
recs = {}
cur_file = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {n: i for i, n in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "" or not r[0].isdigit():
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    st = {}
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            try:
                v = int(r[ix[k]])
            except ValueError:
                v = 0
            if v:
                st[k[6:]] = v
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    if key in recs:
        recs[key][0] += n
        for k, v in st.items():
            recs[key][1][k] = recs[key][1].get(k, 0) + v
    else:
        recs[key] = [n, st]
tot = sum(v[0] for v in recs.values())
print("total samples", tot)
for (f, ln, src), (n, st) in sorted(recs.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{n:6d} {100.0 * n / max(tot, 1):5.1f}%  {f}:{ln:<4d} {src}   {sorted(st.items(), key=lambda x: -x[1])[:3]}")
