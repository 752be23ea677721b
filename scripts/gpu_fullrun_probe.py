"""Whole runs of the reference's e.circle_push bots (to their stop condition): GPU fp64 / fp32 world results beside the frozen
vx3_ref ones (tests/golden/vx3ref_circle_push_*_full.vxbin)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpuvoxels_b200 as gv  # noqa: E402
import oracle_lib as O  # noqa: E402
from golden.make_golden_vx3 import CASES, CIRCLE_PUSH_BOTS  # noqa: E402

models = [gv.Model.from_vxa(*CASES[f"vx3ref_circle_push_{b}_full"][:2]) for b in CIRCLE_PUSH_BOTS]
res = {}
for prec in (gv.FP64, gv.FP32):
    for path in ((0, 1) if prec == gv.FP32 else (0,)):
        ctx = gv.Context(0, prec)
        ctx.set_option("path", path)
        ctx.upload(models)
        ctx.run(1000000, True)
        res[(prec, path)] = ctx.results()
        ctx.close()
for i, b in enumerate(CIRCLE_PUSH_BOTS):
    sc = O.read_vxbin(os.path.join(ROOT, "tests", "golden", f"vx3ref_circle_push_{b}_full.vxbin"))["scalars"][0]
    tr = sc[6:9] - sc[9:12]
    print(f"{b} nv={models[i].num_voxels} ref: steps {int(sc[3])} hits {int(sc[4])} angle {sc[12]:.4f} travel ({tr[0]:+.5f},{tr[1]:+.5f}) fit {sc[5]:+.4f}")
    for k, r in res.items():
        t = np.array(r[i].current_com) - np.array(r[i].initial_com)
        print(f"      fp{k[0]} path{k[1]}: steps {r[i].steps} hits {r[i].collision_count} angle {r[i].recent_angle:.4f} travel ({t[0]:+.5f},{t[1]:+.5f}) fit {r[i].fitness_score:+.4f}"
              f"  |dtravel|/|travel| {np.linalg.norm(t[:2] - tr[:2]) / max(np.linalg.norm(tr[:2]), 1e-12):.2e}")
