#!/usr/bin/env python
"""gpurun_out/allk_<workload>_<tag>.csv (scripts/gpu_allkernels.sh) -> profiles/<tag>_all_kernels.md:
one row per kernel and workload: launches, registers, mean duration, DRAM bytes per launch, achieved DRAM GB/s
(bytes / duration) and that as a fraction of the measured HBM peak (MEASURED_PEAKS.json), SM throughput, issue rate.
usage: summarize_allkernels.py <tag>"""
import collections
import csv
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
peak = 6555.8
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def num(v, unit):
    x = float(v.replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return x * scale.get(unit, 1.0)


out = [f"# Every kernel of the engine under ncu ({tag})", "",
       "`scripts/gpu_allkernels.sh` - `ncu --metrics ... --clock-control none` (one metric pass, cold caches, serialised launches), "
       "short runs of each bench workload.  DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration per launch; "
       f"peak = {peak} GB/s (MEASURED_PEAKS.json).  Kernels whose state lives in shared memory (k_resident) or that are launch-sized "
       "(bookkeeping over a few hundred worlds) are far from the HBM roof by construction.", ""]
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"allk_*_{tag}.csv"))):
    wl = os.path.basename(path)[len("allk_"):-len(f"_{tag}.csv")]
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()  # (kernel, grid, block) -> id -> metric -> value
    for row in csv.DictReader(lines):
        if "Kernel Name" not in row or not row.get("Metric Name"):
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("vx3b::", "")
        key = (name, row["Grid Size"], row["Block Size"])
        try:
            per.setdefault(key, collections.OrderedDict()).setdefault(row["ID"], {})[row["Metric Name"]] = num(row["Metric Value"], row["Metric Unit"])
        except ValueError:
            pass
    # merge launches of one kernel that differ only in grid size
    agg = collections.OrderedDict()
    for (name, grid, block), ids in per.items():
        a = agg.setdefault((name, block), {"n": 0, "grids": [], "rows": []})
        a["n"] += len(ids)
        a["grids"].append(grid)
        a["rows"] += list(ids.values())
    out += [f"## workload `{wl}`", "", "| kernel | block | grids | launches | regs | mean µs | DRAM MB / launch | DRAM GB/s | of HBM peak | SM thr % | issue % | warps active % |",
            "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for (name, block), a in agg.items():
        rows = a["rows"]
        mean = lambda m: sum(r.get(m, 0.0) for r in rows) / max(1, len(rows))
        dur = mean("gpu__time_duration.sum")
        byt = mean("dram__bytes_read.sum") + mean("dram__bytes_write.sum")
        gbs = byt / dur / 1e3 if dur > 0 else 0.0
        grids = sorted(set(a["grids"]))
        gtxt = grids[0] if len(grids) == 1 else f"{grids[0]} … {grids[-1]} ({len(grids)})"
        out.append(f"| `{name}` | {block} | {gtxt} | {a['n']} | {mean('launch__registers_per_thread'):.0f} | {dur:.1f} | {byt / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.3f} | "
                   f"{mean('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {mean('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {mean('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} |")
    out.append("")
dst = os.path.join(ROOT, "profiles", f"{tag}_all_kernels.md")
open(dst, "w").write("\n".join(out) + "\n")
print("wrote", dst)
