"""Ad-hoc GPU parity + timing sweep (run under gpurun while developing; the pytest -m gpu suite is the gate).
Prints one line per (world, precision, path) with the error against the CPU oracle and the device time."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gpuvoxels_b200 as gv  # noqa: E402
import oracle_lib as O  # noqa: E402
from gpuvoxels_b200 import worlds as W  # noqa: E402


def compare(name, model, steps, precisions=(gv.FP64, gv.FP32), paths=(1, 2, 3)):
    ow = O.OracleWorld(model)
    t0 = time.time()
    ow.run(steps)
    tor = time.time() - t0
    ref = ow.state()
    lat = model.array("lattice") * model.desc.vox_size
    disp = np.abs(ref["pos"] - lat).max()
    rr = ow.result()
    for prec in precisions:
        for path in paths:
            try:
                ctx = gv.Context(0, prec)
                ctx.set_option("path", path)
                ctx.set_option("record_pairs", 1)
                ctx.upload([model])
                ctx.run(steps)
                got = ctx.read_state(0)
                ms, launches = ctx.timing()
                r = ctx.results()[0]
                epos = np.abs(got["pos"] - ref["pos"]).max() / max(disp, 1e-30)
                eori = np.abs(got["orient"] - ref["orient"]).max()
                elin = O.rel_err(got["linMom"], ref["linMom"])
                pairs_ok = ""
                if model.desc.enable_collision:
                    gp = ctx.collision_pairs(0)
                    op = ow.collision_pairs()
                    pairs_ok = f" pairs gpu={len(gp)} oracle={len(op)} equal={np.array_equal(gp, op)}"
                print(f"{name:10s} nv={model.num_voxels:6d} steps={steps:5d} fp{prec} path{path}: pos {epos:.2e} orient {eori:.2e} linMom {elin:.2e} "
                      f"com {np.abs(np.array(r.current_com) - np.array(rr.current_com)).max():.2e} fit {r.fitness_score:.6g}/{rr.fitness_score:.6g} "
                      f"hits {r.collision_count}/{rr.collision_count}{pairs_ok} | {ms:.2f} ms {launches} launches, oracle {tor*1e3:.0f} ms", flush=True)
                ctx.close()
            except Exception as e:  # noqa: BLE001
                print(f"{name} fp{prec} path{path}: FAILED {e}", flush=True)


def growth(name, model, checkpoints=(1, 2, 3, 5, 10, 20, 50, 100, 200, 400), prec=None, path=1):
    """Error against the oracle as a function of the step count: separates a bug (wrong from step 1) from amplified rounding."""
    prec = prec or gv.FP64
    ow = O.OracleWorld(model)
    ctx = gv.Context(0, prec)
    ctx.set_option("path", path)
    ctx.upload([model])
    done = 0
    lat = model.array("lattice") * model.desc.vox_size
    for c in checkpoints:
        ow.run(c - done)
        ctx.run(c - done)
        done = c
        got, ref = ctx.read_state(0), ow.state()
        disp = max(np.abs(ref["pos"] - lat).max(), 1e-30)
        e = np.abs(got["pos"] - ref["pos"])
        print(f"growth {name} fp{prec} step {c:4d}: rel pos err xyz {e[:, 0].max() / disp:.2e} {e[:, 1].max() / disp:.2e} {e[:, 2].max() / disp:.2e} (disp {disp:.2e})", flush=True)
    ctx.close()


def main():
    if "--growth" in sys.argv:
        growth("cube5", gv.Model.from_vxa(W.cube_vxa(5)))
        growth("cube5", gv.Model.from_vxa(W.cube_vxa(5)), prec=gv.FP32)
        return
    compare("small", gv.Model.from_vxa(W.small_block_vxa()), 200)
    compare("small", gv.Model.from_vxa(W.small_block_vxa()), 2000)
    compare("small+col", gv.Model.from_vxa(W.small_block_vxa(enable_collision=True)), 500)
    ex = os.path.join(ROOT, "tests", "data", "example.vxa")
    if os.path.exists(ex):
        compare("example", gv.Model.from_vxa(open(ex).read()), 1000)
    compare("xeno0", gv.Model.from_vxa(W.xenobot_base_vxa(), W.xenobot_vxd(0)), 1000)
    compare("xeno1", gv.Model.from_vxa(W.xenobot_base_vxa(), W.xenobot_vxd(1)), 3000, precisions=(gv.FP32,))
    compare("pile", gv.Model.from_vxa(W.contact_pile_vxa()), 600)
    compare("cube12", gv.Model.from_vxa(W.cube_vxa(12)), 300)
    compare("cube30", gv.Model.from_vxa(W.cube_vxa(30)), 100, paths=(1, 3))

    # batch throughput: many xenobot worlds
    for n in (64, 1024):
        base = W.xenobot_base_vxa()
        models = [gv.Model.from_vxa(base, W.xenobot_vxd(i)) for i in range(min(n, 64))]
        models = [models[i % len(models)] for i in range(n)]
        for path in (2, 1):
            for prec in (gv.FP32,):
                ctx = gv.Context(0, prec)
                ctx.set_option("path", path)
                ctx.upload(models)
                steps = 1000 if path == 2 else 200
                ctx.run(steps)
                ms, launches = ctx.timing()
                vs = ctx.total_voxel_steps()
                res = ctx.results()
                rb = np.array([r.contact_rebuilds for r in res]) / steps
                ov = np.array([r.contact_overflows for r in res]) / steps
                print(f"batch n={n} path{path} fp{prec}: {steps} steps {ms:.1f} ms -> {vs / ms * 1e3:.3e} voxel-steps/s ({launches} launches); "
                      f"contact rebuilds/step mean {rb.mean():.3f} max {rb.max():.3f}, overflows/step mean {ov.mean():.4f} max {ov.max():.3f}", flush=True)
                ctx.close()
    # big body streaming throughput
    for n in (50, 100):
        m = gv.Model.from_vxa(W.cube_vxa(n))
        for path in (1, 3):
            ctx = gv.Context(0, gv.FP32)
            ctx.set_option("path", path)
            ctx.upload([m])
            ctx.run(20)
            ctx.run(100)
            ms, launches = ctx.timing()
            print(f"cube{n} path{path} fp32: 100 steps {ms:.1f} ms -> {ctx.total_voxel_steps() / ms * 1e3:.3e} voxel-steps/s", flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
