"""Per-step device time of the cube workload for a list of brick shapes (engine options brick_x/y/z), early and late in the run.
usage: gpu_step_probe.py <cube edge> [bx,by,bz ...]"""
import sys, time
import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(8, 8, 4)]
model = gv.Model.from_vxa(W.cube_vxa(n))
for shape in shapes:
    ctx = gv.Context(0, gv.FP32)
    for k, v in zip("xyz", shape):
        ctx.set_option("brick_" + k, v)
    ctx.upload([model])
    ctx.run(64)
    out = []
    for steps in (64, 512, 1024):
        ctx.run(steps)
        ms, k = ctx.timing()
        out.append(f"{1e3 * ms / steps:6.1f}")
    print(f"brick {shape}: us/step after 64 / 128 / 640 steps: {' '.join(out)}   ({k} launches in the last run)", flush=True)
    ctx.close()
