#!/usr/bin/env python
"""Turns the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.
usage: summarize_profiles.py <tag> [--kernel-regex k_resident]
  gpurun_out/launches_<tag>.csv            -> profiles/<tag>_launches.csv (+ _launches_summary.txt: share of the step per kernel)
  gpurun_out/prof_*_<tag>.ncu-rep          -> profiles/<tag>_<name>_metrics.txt (duration, DRAM bytes, occupancy, issue rate ...)
                                              profiles/<tag>_<name>_lines.txt   (per-source-line instruction / stall shares)
                                              profiles/traffic.json            (dram bytes per launch of each profiled kernel)"""
import collections
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)


def launches():
    for src in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"launches*_{tag}.csv"))):
        launches_one(src, os.path.basename(src)[len("launches"):-len(f"_{tag}.csv")])


def launches_one(src, name):
    shutil.copy(src, os.path.join(OUT, f"{tag}{name}_launches.csv"))
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0] + " grid" + row["Grid Size"] + " block" + row["Block Size"]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values()) or 1
    with open(os.path.join(OUT, f"{tag}{name}_launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none (command: see scripts/gpu_round.sh)\n")
        f.write("per-launch times are serialised and cold-cache: compare SHARES of the step, not absolutes\n\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:100s} launches {n:4d} total {t:12.1f} us  share {100 * t / tot:5.1f}%\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def reports():
    traffic_path = os.path.join(OUT, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_*_{tag}.ncu-rep"))):
        name = os.path.basename(rep)[len("prof_"):-len(f"_{tag}.ncu-rep")]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        with open(os.path.join(OUT, f"{tag}_{name}_metrics.txt"), "w") as f:
            f.write(f"ncu --set full --clock-control none --import-source on  ({os.path.basename(rep)}; command: scripts/gpu_round.sh / gpu_prof.sh)\n")
            for row in rows[2:]:
                d = dict(zip(hdr, row))
                u = dict(zip(hdr, units))
                kname = d.get("Kernel Name", "?")
                f.write(f"\n== {kname}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
                for w in WANT:
                    if w in d:
                        f.write(f"   {w:70s} {d[w]:>18s} {u[w]}\n")
                if "dram__bytes_read.sum" in d:
                    b = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
                    key = kname.split("(")[0].replace("void ", "").replace("vx3b::", "")
                    traffic.setdefault(tag, {}).setdefault(key, []).append({"grid": d.get("Grid Size"), "dram_bytes": b})
        for which in range(2):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, str(which), "40"], capture_output=True, text=True)
            if r.returncode == 0 and r.stdout.strip():
                with open(os.path.join(OUT, f"{tag}_{name}_lines_{which}.txt"), "w") as f:
                    f.write(r.stdout)
    json.dump(traffic, open(traffic_path, "w"), indent=1)


launches()
reports()
print("profiles/:", sorted(os.listdir(OUT)))
