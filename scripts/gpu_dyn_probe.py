"""Ad-hoc probe of the dynamic-topology kernel: runs one golden case and prints progress (unbuffered)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gpuvoxels_b200 as gv
import oracle_lib as O
from golden.make_golden_vx3 import CASES
name = sys.argv[1] if len(sys.argv) > 1 else "vx3ref_sticky_pebbles_3000"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
prec = gv.FP64 if (len(sys.argv) <= 3 or sys.argv[3] == "64") else gv.FP32
print("import ok", flush=True)
vxa, vxd = CASES[name][:2]
m = gv.Model.from_vxa(vxa, vxd)
print("model", m.num_voxels, m.num_links, flush=True)
ctx = gv.Context(0, prec)
print("ctx", flush=True)
ctx.upload([m])
print("upload", flush=True)
t0 = time.time()
ctx.run(steps)
print("run", time.time() - t0, ctx.timing(), flush=True)
r = ctx.results()[0]
print("steps", r.steps, "t", r.current_time, "hits", r.collision_count, "links", ctx.num_links(0), "div", r.diverged, flush=True)
w = O.OracleWorld(m)
w.run(steps)
s, ref = ctx.read_state(0), w.state()
print("pos err", np.abs(s["pos"] - ref["pos"]).max(), "oracle links", len(w.topology()), "oracle hits", w.result().collision_count, flush=True)
sig, fl = ctx.read_signals(0)
osig, ofl = w.signals()
print("sig equal", np.array_equal(sig, osig), np.abs(sig - osig).max(), "topo equal", np.array_equal(ctx.read_topology(0), w.topology()), flush=True)
if len(sys.argv) > 4:  # timing variants: which feature costs what
    import re
    for what, sub in (("all", None), ("no attach", ("<EnableAttach>1", "<EnableAttach>0")), ("no signals", ("<EnableSignals>1", "<EnableSignals>0")),
                      ("no collision, no attach", ("<EnableCollision>1", "<EnableCollision>0"))):
        v = vxa
        if sub:
            v = v.replace(*sub)
            if what.startswith("no collision"):
                v = v.replace("<EnableAttach>1", "<EnableAttach>0")
        mm = gv.Model.from_vxa(v, vxd)
        c = gv.Context(0, prec)
        c.upload([mm])
        c.run(20)
        c.run(steps)
        print(f"{what:28s} {c.timing()[0] / steps * 1e3:9.1f} us per step  (attach {mm.desc.enable_attach} signals {mm.desc.enable_signals} collision {mm.desc.enable_collision})", flush=True)
        c.close()
