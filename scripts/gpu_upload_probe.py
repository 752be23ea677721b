"""Where does vx3b_batch_upload spend its time? (VX3B_DEBUG=1 prints the stages)"""
import sys, time
import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W
which = sys.argv[1] if len(sys.argv) > 1 else "xenobot"
if which == "cube":
    models = [gv.Model.from_vxa(W.cube_vxa(100))]
else:
    base = W.xenobot_base_vxa()
    models = [gv.Model.from_vxa(base, W.xenobot_vxd(i)) for i in range(1024)]
ctx = gv.Context(0, gv.FP32)
for k in range(3):
    t0 = time.perf_counter()
    ctx.upload(models)
    print(f"upload {k}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
t0 = time.perf_counter(); ctx.run(10); print(f"run(10): {1e3 * (time.perf_counter() - t0):.1f} ms")
t0 = time.perf_counter(); r = ctx.results(); print(f"results: {1e3 * (time.perf_counter() - t0):.1f} ms")
