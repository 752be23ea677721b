#!/usr/bin/env python
"""bench.py — BASELINE.json metric: voxel-steps/sec (device-timed) on the xenobot-generation workload.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
  python bench.py --impl reference ...        # the reference's own CPU path (oracle/_ref) on the same worlds

Workload (config.workload): BASELINE.json configs[1] — a generation of 1024 xenobot-scale VXA worlds per GPU (cilia forces,
floor, voxel-voxel contact), each bench "step" = one vx3b_batch_run advancing every world by --sim-steps time steps.
Weak scaling: every rank gets its own 1024 worlds; the only cross-rank traffic is the report gather (fitness per world).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxel-steps/sec"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--worlds", type=int, default=1024, help="worlds per GPU")
    ap.add_argument("--sim-steps", type=int, default=1000, help="time steps every world advances per bench step")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--path", type=int, default=0, help="0 auto (resident + brick/fused), 1 two-kernel streaming, 2 resident, 3 fused, 5 brick")
    ap.add_argument("--workload", default="xenobot", choices=["xenobot", "cube", "slab", "pile"],
                    help="xenobot: BASELINE configs[1] (worlds sharded over the GPUs); cube: configs[2] on one GPU; slab: one --cube^3 body cut into --gpus slabs, halo exchange fused into the step kernel over IPC-mapped peer memory (configs[4], second part; --slab-exchange); pile: configs[3], --pile x --pile bodies of 8^3 voxels (plus a second layer) dropped on each other, voxel-voxel contact on")
    ap.add_argument("--pile", type=int, default=12)
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="engine option (vx3b_ctx_set_option), e.g. contact_skin=0.3")
    ap.add_argument("--verify", action="store_true", help="slab workload: gather the slabs' final state on rank 0 and compare it, bit for bit, with the same body stepped whole on one GPU")
    ap.add_argument("--cube", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-worlds", type=int, default=8)
    ap.add_argument("--workloads", default="all", help="N = 1, default workload only: also measure these other BASELINE configs into a \"workloads\" block "
                    "(all | none | comma list of cube100,pile,xenobot_fp64)")
    ap.add_argument("--slab-cube", type=int, default=160, help="N > 1: edge of the ONE body also stepped as N slabs into a \"slab\" block (0 = skip)")
    ap.add_argument("--first-world", type=int, default=0, help="xenobot workload: index of rank 0's first world (rank r of a multi-GPU run takes worlds first + r * worlds ...)")
    ap.add_argument("--gen-parts", type=int, default=6, help="e2e of a many-world generation: sub-batches of the pipelined runner vx3b_gen_run (1 = plain upload + run + results)")
    ap.add_argument("--slab-sim-steps", type=int, default=200)
    ap.add_argument("--slab-exchange", default="fused", choices=["fused", "device", "nccl"],
                    help="slab halo exchange: fused = the step kernel stores boundary voxels straight into the neighbour's IPC-mapped state arrays and raises its flag "
                         "(one launch per time step); device = separate push / pull kernels through a mapped receive area; nccl = round 1's per-step ncclSend/Recv")
    ap.add_argument("--slab-big", type=int, default=400, help="N = 8: edge of the big body (BASELINE configs[4]: 400^3) also stepped as 8 slabs into a \"slab_big\" block (0 = skip)")
    return ap.parse_args()


def workload_config(args):
    if args.workload == "slab":
        return {"workload": f"BASELINE configs[4] (second part): ONE {args.cube}^3-voxel soft body under gravity with floor contact, split into {args.gpus} z-slab(s)",
                "worlds_per_gpu": 1, "sim_steps_per_step": args.sim_steps,
                "parallelism": f"{args.gpus} slab(s), one per GPU; per time step the boundary layer (pos, temperature, orientation) goes to each neighbour "
                               + {"fused": "by stores of the step kernel itself into its IPC-mapped state arrays over NVLink", "device": "by a push kernel's stores into its IPC-mapped receive area over NVLink",
                                  "nccl": "by one NCCL send/recv"}[getattr(args, "slab_exchange", "fused")],
                "l2": "slab state larger than L2"}
    if args.workload == "pile":
        return {"workload": f"BASELINE configs[3]: multi-body scene, {args.pile}x{args.pile} + {args.pile - 1}x{args.pile - 1} bodies of 8^3 voxels piling up, voxel-voxel self-collision enabled (grid broad phase + candidate lists)",
                "worlds_per_gpu": 1, "sim_steps_per_step": args.sim_steps, "parallelism": "replicas", "l2": "explicit L2 flush between steps"}
    if args.workload == "cube":
        return {"workload": f"BASELINE configs[2]: single {args.cube}^3-voxel soft body under gravity with floor contact, no self-collision",
                "worlds_per_gpu": 1, "sim_steps_per_step": args.sim_steps, "parallelism": "replicas", "l2": "state larger than L2 + explicit L2 flush between steps"}
    return {"workload": "BASELINE configs[1]: xenobot generation, 1024 small VXA worlds (~500 voxels, cilia forces, floor, contact) batched per B200",
            "worlds_per_gpu": args.worlds, "sim_steps_per_step": args.sim_steps, "parallelism": f"worlds sharded over {args.gpus} GPU(s), no data-path collective",
            "l2": "explicit L2 flush (256 MiB write) between timed steps"}


def build_worlds(args, rank):
    import gpuvoxels_b200 as gv
    texts = world_texts(args, rank)
    return [gv.Model.from_vxa(a, d) for a, d in texts], texts


def algorithmic_bytes_per_sim_step(models, precision):
    # fp32 112*Nv + 88*Nl, fp64 216*Nv + 160*Nl bytes per time step
    nv = sum(m.num_voxels for m in models)
    nl = sum(m.num_links for m in models)
    return (112 * nv + 88 * nl) if precision == 32 else (216 * nv + 160 * nl)


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- reference (CPU) arm
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "vx2_ref")
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


def run_reference_sample(texts, sim_steps, procs, collisions=True, threads=1):
    """Runs the UNMODIFIED reference CPU library on a sample of the same worlds, `procs` single-threaded processes at a time
    (the reference has no CPU batch mode; one process per core is how a generation uses a whole host).  Returns
    (voxel_steps_per_sec, total_voxel_steps, wall_seconds)."""
    from gpuvoxels_b200.worlds import merge_vxd  # pure Python (xml.etree): this arm never loads libvx3b200.so
    ref = reference_binary()
    with tempfile.TemporaryDirectory() as td:
        jobs = []
        for i, (vxa, vxd) in enumerate(texts):
            p = os.path.join(td, f"w{i}.vxa")
            with open(p, "w") as f:
                f.write(merge_vxd(vxa, vxd))
            cmd = [ref, "--vxa", p, "--steps", str(sim_steps), "--vx3", "--enable-cilia", "1", "--threads", str(threads), "--time"]
            if collisions:
                cmd += ["--vx2-collisions", "1"]
            jobs.append(cmd)
        t0 = time.perf_counter()
        running, outs, it = [], [], iter(jobs)
        pending = True
        while pending or running:
            while pending and len(running) < procs:
                try:
                    running.append(subprocess.Popen(next(it), stdout=subprocess.PIPE, text=True))
                except StopIteration:
                    pending = False
            for p in list(running):
                if p.poll() is not None:
                    outs.append(p.stdout.read())
                    running.remove(p)
            time.sleep(0.002)
        wall = time.perf_counter() - t0
    vs, rates = 0, []
    for o in outs:
        j = json.loads(o.strip().splitlines()[-1])
        vs += j["voxels"] * j["steps"]
        rates.append(j["voxel_steps_per_s"])  # steady_clock around the step loop only (parse/Import excluded)
    return min(procs, len(jobs)) * float(np.mean(rates)), vs, wall


def world_texts(args, rank):
    """(base VXA, VXD) text of every world of this rank - no host model, no library"""
    from gpuvoxels_b200 import worlds as W
    if args.workload in ("cube", "slab"):
        return [(W.cube_vxa(args.cube), None)]
    if args.workload == "pile":
        return [(W.contact_pile_vxa(k=args.pile, cube=8), None)]
    base = W.xenobot_base_vxa()
    world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
    if world == 1:
        return [(base, W.xenobot_vxd(args.first_world + i)) for i in range(args.worlds)]
    # a generation over several GPUs is dealt by cost, longest-processing-time-first, as vx3b_job_run deals a directory over its devices
    # (vx3b_shard_worlds; every rank computes the same deal): contiguous index ranges differ by 9 % in run time (worlds 1,024-2,047
    # take 92.8 ms where worlds 0-1,023 take 85.2), which a max over ranks then reports as lost scaling
    import ctypes as C
    import gpuvoxels_b200 as gv
    n = world * args.worlds
    cost = (C.c_int64 * n)(*[W.xenobot_cost(args.first_world + i) for i in range(n)])
    shard = (C.c_int32 * n)()
    if gv.lib().vx3b_shard_worlds(cost, n, world, shard) != 0:
        raise RuntimeError("vx3b_shard_worlds failed")
    return [(base, W.xenobot_vxd(args.first_world + i)) for i in range(n) if shard[i] == rank]


def oracle_port_sample(models, sim_steps):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    t0 = time.perf_counter()
    vs = 0
    for m in models:
        w = O.OracleWorld(m)
        w.run(sim_steps)
        vs += m.num_voxels * sim_steps
    wall = time.perf_counter() - t0
    return vs / wall, vs, wall


def cpu_baseline(args, models, texts):
    cores = os.cpu_count() or 1
    # xenobot: enough worlds for several seconds on every host core (32 per core); the big single-body workloads have one world
    n = min(len(texts), max(args.cpu_sample_worlds, 32 * cores))
    steps = min(args.sim_steps, 1000 if args.workload == "xenobot" else 10 if args.workload == "pile" else 5)
    if reference_binary():
        col = args.workload in ("xenobot", "pile")
        if len(texts) == 1:
            # one body: the reference's OpenMP loops over links / voxels (USE_OMP), timed at 1 thread and at every host core
            v1, vs, wall1 = run_reference_sample(texts, steps, 1, collisions=col, threads=1)
            vn, _, walln = run_reference_sample(texts, steps, 1, collisions=col, threads=cores)
            best, used = (vn, cores) if vn > v1 else (v1, 1)
            return {"value": best, "unit": METRIC, "cores": used, "kind": "reference", "value_1_thread": v1, f"value_{cores}_threads": vn,
                    "sample": f"the same body x {steps} time steps, oracle/_ref/vx2_ref (unmodified CVoxelyze, OpenMP over links and voxels"
                              f"{', its own collision system' if col else ''}) at OMP threads = 1 and = {cores}; step loop only; {vs} voxel-steps, "
                              f"{wall1:.1f} s + {walln:.1f} s of wall clock including parse"}
        procs = min(cores, n)
        v, vs, wall = run_reference_sample(texts[:n], steps, procs, collisions=col)
        return {"value": v, "unit": METRIC, "cores": procs, "kind": "reference",
                "sample": f"{n} of the {len(texts)} worlds x {steps} time steps, oracle/_ref/vx2_ref (unmodified CVoxelyze + VX3 stepping deltas, its own "
                          f"collision system), {procs} single-threaded processes in parallel; {vs} voxel-steps in {wall:.1f} s"}
    v, vs, wall = oracle_port_sample(models[:n], steps)
    return {"value": v, "unit": METRIC, "cores": 1, "kind": "port",
            "sample": f"{n} of the {len(texts)} worlds x {steps} time steps, scalar CPU restatement (oracle/vx3_oracle.cpp), 1 thread; {vs} voxel-steps in {wall:.1f} s"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    texts = world_texts(args, 0)  # text only: this arm does not load the product library
    models = None
    cores = os.cpu_count() or 1
    n = min(len(texts), max(8 * cores, args.cpu_sample_worlds))
    steps = min(args.sim_steps, 500 if args.workload == "xenobot" else 3)
    times, vs_total, rates = [], 0, []
    kind = "reference" if reference_binary() else "port"
    for it in range(args.warmup + args.steps):
        if kind == "reference":
            one = len(texts) == 1  # a single body uses the reference's OpenMP threads, a generation one process per core
            v, vs, wall = run_reference_sample(texts[:n], steps, min(cores, n), collisions=args.workload in ("xenobot", "pile"), threads=cores if one else 1)
        else:
            if models is None:
                models = build_worlds(args, 0)[0]
            v, vs, wall = oracle_port_sample(models[:n], steps)
        if it >= args.warmup:
            times.append(wall)
            vs_total += vs
            rates.append(v)
    value = float(np.mean(rates))
    used = (cores if len(texts) == 1 else min(cores, n)) if kind == "reference" else 1
    sample = (f"each step = {n} of the {len(texts)} worlds x {steps} time steps on the host CPU "
              f"({'oracle/_ref/vx2_ref, unmodified reference CPU library' if kind == 'reference' else 'oracle port'}, {used} core(s))")
    cfg = workload_config(args)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": cfg, "cpu_baseline": {"value": value, "unit": METRIC, "cores": used, "kind": kind, "sample": sample},
                      "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
    return 0


# ------------------------------------------------------------------------------------------------- our arm
def upload_bytes(models):
    import ctypes as C
    from gpuvoxels_b200 import _abi
    total = 0
    for m in models:
        d = m.desc
        nv, nl = d.num_voxels, d.num_links
        total += C.sizeof(_abi.WorldDesc) + nv * (6 + 4 + 8 + 24 + 24 + 1 + 12 + 12 + 24 + 24) + nl * (4 + 4 + 1 + 4)
        total += d.num_voxel_mats * C.sizeof(_abi.VoxelMaterial) + d.num_link_mats * C.sizeof(_abi.LinkMaterial)
    return total


def measure(args, rank, world, local):
    """One workload on this rank's GPU: device-timed throughput with the worlds resident in HBM, the end-to-end number through the
    C ABI with host buffers, roofline of the dominant kernel, CPU baseline (rank 0, N = 1).  Returns the JSON object on rank 0."""
    import torch
    import torch.distributed as dist
    import ctypes as C
    import gpuvoxels_b200 as gv
    from gpuvoxels_b200 import _abi

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    models, texts = build_worlds(args, rank)
    ctx = gv.Context(local, args.precision)
    ctx.set_option("path", args.path)
    for o in args.opt:
        ctx.set_option(o.split("=")[0], float(o.split("=")[1]))
    ctx.upload(models)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    bytes_per_sim_step = algorithmic_bytes_per_sim_step(models, args.precision)

    # ---- device-timed throughput: state resident in HBM, K steps -------------------------------------------------
    for _ in range(args.warmup):
        ctx.run(args.sim_steps)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    dev_ms, vsteps, launches = 0.0, 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.run(args.sim_steps)
        ms, k = ctx.timing()
        dev_ms += ms
        launches += k
        vsteps += ctx.total_voxel_steps()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    results = ctx.results()
    diverged = sum(r.diverged for r in results)

    # ---- end to end through the C ABI with host buffers: upload + run + results (+ report gather) ----------------------
    h2d = upload_bytes(models)
    d2h = len(models) * C.sizeof(_abi.WorldResult)
    # a generation of many worlds goes through the pipelined runner (vx3b_gen_run: sub-batch k + 1 is staged and copied while sub-batch k
    # already steps); one big world through the plain context calls
    pipelined = len(models) >= 16 and args.gen_parts > 1
    ctx2 = gv.Generation(local, args.precision, args.gen_parts) if pipelined else gv.Context(local, args.precision)
    ctx2.set_option("path", args.path)
    for o in args.opt:
        ctx2.set_option(o.split("=")[0], float(o.split("=")[1]))
    e2e_steps = max(1, min(args.steps, 3))

    def generation():
        if pipelined:
            return ctx2.run(models, args.sim_steps)
        ctx2.upload(models)
        ctx2.run(args.sim_steps)
        return ctx2.results()

    generation()
    barrier()
    t1 = time.perf_counter()
    e2e_vs = 0
    for _ in range(e2e_steps):
        res = generation()
        fit = torch.tensor([r.fitness_score for r in res], dtype=torch.float64, device="cuda")
        if world > 1:
            gathered = [torch.empty_like(fit) for _ in range(world)]
            dist.all_gather(gathered, fit)
        _ = fit.cpu()
        e2e_vs += ctx2.total_voxel_steps()
    barrier()
    e2e_wall = time.perf_counter() - t1
    ctx2.close()

    # ---- reduce over ranks: time = max, work = sum -----------------------------------------------------------------------
    stats = torch.tensor([dev_ms, wall, e2e_wall], dtype=torch.float64, device="cuda")
    work = torch.tensor([float(vsteps), float(e2e_vs), float(launches), float(diverged)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_max, e2e_wall_max = stats.tolist()
    vs_all, e2e_vs_all, launches_all, diverged_all = work.tolist()

    out = None
    if rank == 0:
        value = vs_all / (dev_ms_max * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "of fallback (B200_PROFILING.md)"
        # dominant kernel: the resident (or streaming link+voxel) step kernel; algorithmic bytes per launch = per-step bytes x steps
        achieved = bytes_per_sim_step * args.sim_steps / ((dev_ms / args.steps) * 1e-3) / 1e9
        if args.path == 2 or (args.path == 0 and args.workload == "xenobot"):
            kernel = "k_resident"  # one launch per size class steps every world through all sim_steps
        elif args.workload == "pile":
            kernel = "k_fused"     # + k_cgrid_insert / k_cgrid_narrow (contact) + k_world_step: four launches per time step
        elif args.path in (0, 5):
            kernel = "k_brick"     # one launch per time step
        else:
            kernel = "k_fused" if args.path == 3 else "k_links+k_voxels"
        # DRAM bytes of the dominant kernel from the newest committed `ncu --set full` capture (profiles/traffic.json, written by
        # scripts/summarize_profiles.py): for k_brick per launch = per time step; for k_resident summed over the size-class
        # launches of the captured run (state is loaded and stored once per launch, so it does not scale with sim_steps)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            for tag in sorted(tj):
                rows = [r for k, v in tj[tag].items() if k.startswith(kernel) for r in v]
                if rows:
                    per_launch = [r["dram_bytes"] for r in rows]
                    traffic = sum(per_launch) / len(per_launch) if kernel == "k_brick" else sum(per_launch)
                    traffic_src = f"profiles/traffic.json[{tag}]"
        out = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 arithmetic, f64 positions" if args.precision == 32 else "f64", "data": "synthetic", "config": workload_config(args),
            "voxels_per_gpu": sum(m.num_voxels for m in models), "links_per_gpu": sum(m.num_links for m in models),
            "wall_s_timed_region": wall_max, "diverged_worlds": int(diverged_all),
            "contact_rebuilds_per_world_rank0": float(np.mean([r.contact_rebuilds for r in results])), "contact_overflows_rank0": int(sum(r.contact_overflows for r in results)),
            "steps_per_world_rank0": float(np.mean([r.steps for r in results])), "engine_options": args.opt,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": kernel, "peak_source": peak_src, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": bytes_per_sim_step * args.sim_steps,
                         "note": "algorithmic bytes = (112*Nv + 88*Nl) per time step (SURVEY.md 8d) x time steps per launch, rank 0's worlds"},
            "e2e": {"value": e2e_vs_all / e2e_wall_max, "unit": METRIC, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": (f"vx3b_gen_run (host arrays -> HBM, run, results; {args.gen_parts} sub-batches, each stepping while the next is uploaded)" if pipelined
                             else "vx3b_batch_upload (host arrays -> HBM) + vx3b_batch_run + vx3b_batch_results") + " (+ NCCL gather of fitness for N>1), wall clock"},
            "gpu_launches": int(launches_all), "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args, models, texts)
    ctx.close()
    return out if rank == 0 else None


# the other single-GPU configurations of BASELINE.json, measured in the same run so the driver's record carries them:
# name -> (argument overrides, bench steps, warm-up steps)
EXTRA_WORKLOADS = {
    "cube100": (dict(workload="cube", cube=100, sim_steps=200, precision=32), 3, 3),
    "pile": (dict(workload="pile", pile=12, sim_steps=200, precision=32), 3, 3),
    "xenobot_fp64": (dict(workload="xenobot", sim_steps=200, precision=64), 2, 3),
}


def main_ours(args):
    import copy
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = measure(args, rank, world, local)
    names = [] if args.workloads == "none" else list(EXTRA_WORKLOADS) if args.workloads == "all" else args.workloads.split(",")
    if world == 1 and args.workload == "xenobot" and args.precision == 32 and names:
        out["workloads"] = {}
        for name in names:
            over, k, w = EXTRA_WORKLOADS[name]
            a2 = copy.copy(args)
            for key, v in over.items():
                setattr(a2, key, v)
            a2.steps, a2.warmup = k, w
            try:
                r = measure(a2, rank, world, local)
                keep = ("value", "unit", "steps", "warmup", "ms_per_step", "dtype", "config", "voxels_per_gpu", "links_per_gpu", "diverged_worlds", "roofline", "e2e",
                        "gpu_launches", "clocks", "cpu_baseline", "contact_rebuilds_per_world_rank0")
                out["workloads"][name] = {key: r[key] for key in keep if key in r}
            except Exception as e:  # noqa: BLE001 - the headline line must still be printed
                out["workloads"][name] = {"error": f"{type(e).__name__}: {e}"}
    if world > 1:
        # BASELINE configs[4], second part, in the same multi-GPU run: ONE body cut into N slabs (strong scaling); at N = 8 also the
        # 400^3 body the config names
        for key, cube, sim_steps in (("slab", args.slab_cube, args.slab_sim_steps), ("slab_big", args.slab_big if world >= 8 else 0, 100)):
            if cube <= 0:
                continue
            a2 = copy.copy(args)
            a2.workload, a2.cube, a2.sim_steps, a2.steps, a2.warmup, a2.verify = "slab", cube, sim_steps, 3, 3, False
            try:
                slab = measure_slab(a2, rank, world, local)
            except Exception as e:  # noqa: BLE001
                slab = {"error": f"{type(e).__name__}: {e}"}
            if rank == 0:
                out[key] = slab
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main_slab(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = measure_slab(args, rank, world, local)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def measure_slab(args, rank, world, local):
    """One body over N GPUs (strong scaling: the body is fixed, N slabs).  A bench step = --sim-steps time steps, each followed by
    the halo exchange; device time from CUDA events on the slab's stream, max over ranks."""
    import torch
    import torch.distributed as dist
    import gpuvoxels_b200 as gv
    from gpuvoxels_b200 import slab as S
    from gpuvoxels_b200 import worlds as W

    t_host = time.perf_counter()
    vxa = W.cube_vxa(args.cube)
    t_text = time.perf_counter() - t_host
    device_exchange = args.slab_exchange in ("fused", "device")
    # every rank builds ONLY its own slab, in C++ (vx3b_model_from_vxa_slab); the whole body never exists on any rank
    part = S.native_part(vxa, None, rank, world)
    dt = part.body_dt
    t_split = time.perf_counter() - t_host - t_text
    body_voxels, body_links = int(part.body_voxels), int(part.body_links)
    run = S.SlabRunner(part, local, args.precision, dt)
    t_upload = time.perf_counter() - t_host - t_text - t_split
    b = run.buf
    if device_exchange and world > 1:
        # receive areas are mapped into the neighbours' processes (CUDA IPC, stores over NVLink): the handles travel once, here
        mine = run.open_endpoints()
        eps = [None] * world
        dist.all_gather_object(eps, mine)
        run.connect(eps[rank - 1]["hi"] if rank > 0 else None, eps[rank + 1]["lo"] if rank < world - 1 else None, fused=args.slab_exchange == "fused")
        dist.barrier()

    def exchange():
        if world == 1:
            return
        ops = []
        if rank > 0:
            ops += [dist.P2POp(dist.isend, b[S.SEND_LO], rank - 1), dist.P2POp(dist.irecv, b[S.GHOST_LO], rank - 1)]
        if rank < world - 1:
            ops += [dist.P2POp(dist.isend, b[S.SEND_HI], rank + 1), dist.P2POp(dist.irecv, b[S.GHOST_HI], rank + 1)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()  # stream-ordered: the slab's stream waits for the transfer, the host does not

    def steps(n):
        if device_exchange:
            run.run(n)  # step kernel + push + pull per time step, 64 time steps per CUDA graph, nothing on the host in between
            return
        with torch.cuda.stream(run.stream):
            for _ in range(n):
                run.step_begin()
                exchange()
                run.step_end()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        steps(args.sim_steps)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t0 = time.perf_counter()
    with torch.cuda.stream(run.stream):
        ev[0].record()
    for _ in range(args.steps):
        steps(args.sim_steps)
    with torch.cuda.stream(run.stream):
        ev[1].record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev[0].elapsed_time(ev[1])
    clocks = sampler.stop() if sampler else None
    res = run.ctx.results()[0]
    timeouts = run.ctx.halo_timeouts()
    wait_us_per_step = run.ctx.halo_wait_ns() / 1e3 / max(1, args.sim_steps * (args.warmup + args.steps)) / (2 if 0 < rank < world - 1 else 1)
    waits = [None] * world
    if world > 1:
        dist.all_gather_object(waits, round(wait_us_per_step, 1))
    # whole-body centre of mass across the slabs (one all-reduce of <= 8 doubles): every slab reports the centre of
    # mass of its own voxels (ghosts excluded); the body here is of one material, so the weights are the own-voxel counts
    com = torch.tensor([c * part.num_own for c in res.current_com] + [float(part.num_own)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(com, op=dist.ReduceOp.SUM)
    body_com = (com[:3] / com[3]).tolist()
    verified = None
    if args.verify:  # parity of the sliced run over real NCCL: every voxel of every slab against the whole body on rank 0's GPU
        state, gidx = run.own_state()
        parts_state = [None] * world
        if world > 1:
            dist.all_gather_object(parts_state, (gidx, state))
        else:
            parts_state = [(gidx, state)]
        if rank == 0:
            total = args.sim_steps * (args.warmup + args.steps)
            model = gv.Model.from_vxa(vxa)
            whole = gv.Context(local, args.precision)
            whole.set_option("path", 5)
            whole.upload([model])
            whole.run(total)
            ref = whole.read_state(0)
            whole.close()
            verified = all(np.array_equal(st[k], ref[k][g]) for g, st in parts_state for k in ("pos", "orient", "linMom", "angMom"))
    own = part.num_own
    nl_own = int(part.model.num_links)
    stats = torch.tensor([dev_ms, wall], dtype=torch.float64, device="cuda")
    work = torch.tensor([float(own) * args.sim_steps * args.steps, float(res.diverged), float(run.halo_bytes_per_step), float(timeouts)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_max = stats.tolist()
    vs_all, diverged_all, halo_all, timeouts_all = work.tolist()
    out = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "of fallback (B200_PROFILING.md)"
        nsteps = args.sim_steps * args.steps
        alg = (112 * body_voxels + 88 * body_links) if args.precision == 32 else (216 * body_voxels + 160 * body_links)
        achieved = alg * nsteps / (dev_ms_max * 1e-3) / 1e9 / world  # per GPU
        out = {"metric": METRIC, "value": vs_all / (dev_ms_max * 1e-3), "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32 arithmetic, f64 positions" if args.precision == 32 else "f64", "data": "synthetic", "config": workload_config(args),
               "voxels_total": body_voxels, "links_total": body_links, "voxels_rank0_incl_ghosts": part.model.num_voxels, "links_rank0": nl_own,
               "exchange": "fused into k_brick: boundary voxels stored straight into the neighbour's IPC-mapped state arrays over NVLink, flag raised by the last CTA, next launch waits for it; 64 time steps per CUDA graph" if args.slab_exchange == "fused"
               else "device-side: k_halo_push into the neighbour's IPC-mapped receive area + flag, k_halo_pull; 64 time steps per CUDA graph" if device_exchange
               else "host-driven: k_halo_export, ncclSend/ncclRecv (torch.distributed.batch_isend_irecv), k_halo_import per time step",
               "halo_pull_timeouts": int(timeouts_all), "halo_wait_us_per_time_step_by_rank": waits, "body_center_of_mass_allreduced": body_com,
               "host_s_rank0": {"vxa_text": t_text, "slab_model_cpp": t_split, "upload": t_upload},
               "us_per_time_step": 1e3 * dev_ms_max / nsteps, "halo_bytes_per_time_step_all_ranks": int(halo_all), "wall_s_timed_region": wall_max,
               "diverged_worlds": int(diverged_all), "verified_bit_exact_against_whole_body": verified,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s per GPU", "frac": achieved / peak, "traffic": None, "kernel": "k_brick",
                            "peak_source": peak_src, "note": "algorithmic bytes of the WHOLE body (112*Nv + 88*Nl per time step) / N GPUs / device time"},
               "e2e": {"value": vs_all / wall_max, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "what": "same timed region by host wall clock (graph launches, or per-step launches + NCCL calls, included); the body is uploaded once before it"},
               "gpu_launches": int(nsteps * (1 if args.slab_exchange == "fused" or world == 1 else 3 if device_exchange else 5)), "clocks": clocks}
        if timeouts_all:
            out["error"] = "a halo pull timed out: the run is invalid"
    run.close()
    return out if rank == 0 else None


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_slab(a) if a.workload == "slab" else main_ours(a))
