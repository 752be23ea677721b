#!/usr/bin/env python
"""Dynamic instruction / stall-sample profile of k_resident PER PHASE from an `ncu --set full --import-source on` capture of the
RES_MARKERS build (pmevent markers = PMTRIG in SASS at the phase boundaries): sums the per-SASS-instruction counters of the source
page between consecutive markers, in address order.
usage: ncu_phases.py report.ncu-rep [kernel-index]"""
import csv
import re
import subprocess
import sys
from collections import Counter, defaultdict

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels, cur, hdr = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name" or (len(r) > 1 and r[0] == "Kernel Name"):
        cur = {"name": r[1], "rows": []}
        kernels.append(cur)
        hdr = None
    elif r[0] in ("Address", "#") or (hdr is None and "Source" in r and "Instructions Executed" in r):
        hdr = r
        if cur is not None:
            cur["hdr"] = r
    elif cur is not None and hdr is not None and len(r) == len(hdr):
        cur["rows"].append(r)
kernels = [k for k in kernels if k["rows"]]
if not kernels:
    print("no per-instruction rows found; first lines:\n" + "\n".join(out.splitlines()[:15]))
    sys.exit(1)
k = kernels[min(which, len(kernels) - 1)]
h = k["hdr"]
ix = {n: i for i, n in reversed(list(enumerate(h)))}
src = ix.get("Source")
inst = ix.get("Instructions Executed")
smp = ix.get("# Samples", ix.get("Samples"))
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
names = {0: "before the loop (load)", 1: "rebuild", 2: "link pass", 3: "narrow phase", 4: "barrier 1", 5: "voxel pass", 6: "bookkeeping + barrier 2", 7: "loop end / write-back"}
phase = 0
ci, cs, ops, st = Counter(), Counter(), defaultdict(Counter), defaultdict(Counter)
for r in k["rows"]:
    text = r[src]
    m = re.search(r"PMTRIG\s+(0x[0-9a-f]+|\d+)", text)
    if m:
        phase = int(m.group(1), 0).bit_length() - 1 if int(m.group(1), 0) > 0 else 0
    n = int(float(r[inst] or 0))
    ci[phase] += n
    cs[phase] += int(float(r[smp] or 0)) if smp is not None else 0
    op = re.sub(r"^@!?U?P\d+\s+", "", text.strip()).split(" ")[0].split(".")[0]
    ops[phase][op] += n
    for s in stalls:
        st[phase][s[6:]] += int(float(r[ix[s]] or 0))
ti, ts = sum(ci.values()) or 1, sum(cs.values()) or 1
print("kernel:", k["name"][:110])
print(f"warp instructions {ti}, samples {ts}")
for p in sorted(ci):
    top = ", ".join(f"{o} {100 * c / max(ci[p], 1):.0f}%" for o, c in ops[p].most_common(10))
    sr = ", ".join(f"{o} {100 * c / max(cs[p], 1):.0f}%" for o, c in st[p].most_common(5))
    print(f"{names.get(p, p):28s} inst {100 * ci[p] / ti:5.1f}%  samples {100 * cs[p] / ts:5.1f}%\n      ops: {top}\n      stalls: {sr}")
