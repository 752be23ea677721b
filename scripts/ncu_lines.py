#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [kernel-index] [top-N]   -> file:line, % instructions, % samples, top stall reasons"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split into (file, function) sections
sections, cur, fpath = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
    elif r[0] == "Function Name":
        cur = {"file": fpath, "func": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]) and r[0] != "":
        cur["rows"].append(r)
funcs = []
for s in sections:
    if s["func"] not in funcs:
        funcs.append(s["func"])
func = funcs[which]
print("kernel:", func[:100])
lines = []
for s in sections:
    if s["func"] != func:
        continue
    h = s["hdr"]
    idx = {}
    for i, n in enumerate(h):
        idx.setdefault(n, i)
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    for r in s["rows"]:
        inst = int(r[idx["Instructions Executed"]] or 0)
        smp = int(r[idx["# Samples"]] or 0)
        st = {n: int(r[idx[n]] or 0) for n in stalls}
        lines.append((s["file"].split("/")[-1], int(r[0]), r[1].strip()[:90], inst, smp, st))
ti = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print(f"total warp instructions {ti}, samples {ts}")
agg = defaultdict(int)
for l in lines:
    for k, v in l[5].items():
        agg[k] += v
print("stalls:", ", ".join(f"{k[6:]} {100 * v / ts:.0f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    st = ", ".join(f"{k[6:]} {v}" for k, v in sorted(l[5].items(), key=lambda x: -x[1])[:3] if v)
    print(f"{l[0]}:{l[1]:4d} inst {100 * l[3] / ti:5.1f}% samples {100 * l[4] / ts:5.1f}%  [{st}]  {l[2]}")
