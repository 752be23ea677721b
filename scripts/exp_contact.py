"""Experiment: resident-kernel throughput and contact-list rebuild rate vs skin / pebbles / threads."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W

def run(models, steps, **opts):
    ctx = gv.Context(0, gv.FP32)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.upload(models)
    ctx.run(steps)
    ctx.run(steps)
    ms, launches = ctx.timing()
    res = ctx.results()
    rb = np.array([r.contact_rebuilds for r in res]) / (2 * steps)
    ov = np.array([r.contact_overflows for r in res]) / (2 * steps)
    vs = ctx.total_voxel_steps()
    ctx.close()
    return f"{ms:8.1f} ms {vs / ms * 1e3:.3e} vs/s rebuild/step {rb.mean():.3f} (max {rb.max():.3f}) overflow {ov.mean():.4f}"

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
base = W.xenobot_base_vxa()
uniq = 256
for peb in (True,):
    models = [gv.Model.from_vxa(base, W.xenobot_vxd(i, pebbles=peb)) for i in range(uniq)]
    models = [models[i % uniq] for i in range(n)]
    for skin in (0.1, 0.15, 0.25):
        print(f"pebbles={peb} skin={skin}:", run(models, 1000, contact_skin=skin), flush=True)
base_nc = W.xenobot_base_vxa(enable_collision=False)
models = [gv.Model.from_vxa(base_nc, W.xenobot_vxd(i)) for i in range(uniq)]
models = [models[i % uniq] for i in range(n)]
print("no collision:", run(models, 1000), flush=True)
for t in ():
    print(f"no collision threads={t}:", run(models, 1000, resident_threads=t), flush=True)
