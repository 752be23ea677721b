#!/usr/bin/env python
"""SASS opcode histogram of the first kernel in an ncu report (warp-level instructions executed).
usage: ncu_ops.py report.ncu-rep [top-N]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; ops = defaultdict(int); tot = 0; kernels = 0
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1: break
        print(r[1][:100]); continue
    if r[0] == "Address": hdr = r; continue
    if hdr and len(r) >= 6:
        inst = int(r[5] or 0)
        src = r[1].strip().split()
        if not src: continue
        op = src[1] if src[0].startswith("@") else src[0]
        op = op.split(".")[0]
        ops[op] += inst; tot += inst
print("total warp instructions", tot)
for k, v in sorted(ops.items(), key=lambda x: -x[1])[:top]:
    print(f"{k:10s} {v:12d} {100*v/tot:5.1f}%")
