"""Per-step time of the cube workload as a function of run length and graph chunk (where does the time between kernels go?)."""
import sys, time
import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
model = gv.Model.from_vxa(W.cube_vxa(n))
for chunk in (64, 512):
    ctx = gv.Context(0, gv.FP32)
    ctx.set_option("chunk", chunk)
    ctx.upload([model])
    ctx.run(64)
    for steps in (64, 512, 512, 2048):
        t0 = time.perf_counter()
        ctx.run(steps)
        wall = time.perf_counter() - t0
        ms, k = ctx.timing()
        print(f"chunk {chunk:4d} steps {steps:5d}: device {1e3 * ms / steps:7.1f} us/step, wall {1e6 * wall / steps:7.1f} us/step, launches {k}, total steps so far {ctx.results()[0].steps}")
    ctx.close()
