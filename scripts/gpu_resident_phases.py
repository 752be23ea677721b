"""Where k_resident's time goes, phase by phase (needs the RES_PROFILE build:
   VX3B_VARIANT=prof VX3B_NVCC_EXTRA=-DRES_PROFILE python __graft_entry__.py ; VX3B_LIB=gpuvoxels_b200/libvx3b200_prof.so python scripts/gpu_resident_phases.py)
Prints warp-cycles per phase as a share of the total and per world-step."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuvoxels_b200 as gv  # noqa: E402
from gpuvoxels_b200 import worlds as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
base = W.xenobot_base_vxa()
models = [gv.Model.from_vxa(base, W.xenobot_vxd(i)) for i in range(n)]
lib = gv.lib()
lib.vx3b_debug_resident_profile.argtypes = [C.POINTER(C.c_uint64), C.c_int]
ctx = gv.Context(0, gv.FP32)
for o in sys.argv[3:]:
    ctx.set_option(o.split("=")[0], float(o.split("=")[1]))
ctx.upload(models)
ctx.run(steps)
buf = (C.c_uint64 * 16)()
lib.vx3b_debug_resident_profile(buf, 1)
ctx.run(steps)
ms, _ = ctx.timing()
lib.vx3b_debug_resident_profile(buf, 1)
v = list(buf)
names = ["rebuild (+ its barriers)", "link pass", "narrow phase", "wait at barrier 1", "voxel pass", "bookkeeping + barrier 2"]
tot = sum(v[:6])
print(f"{n} worlds x {steps} steps: {ms:.1f} ms -> {ctx.total_voxel_steps() / ms * 1e3:.3e} voxel-steps/s; warp-steps {v[8]}, warp-rebuilds {v[9]} ({v[9] / max(v[8], 1):.3f} per step)")
print(f"link rounds per step {v[10] / max(v[8], 1):.2f}, voxel rounds per step {v[11] / max(v[8], 1):.2f}")
for i, nm in enumerate(names):
    print(f"  {nm:28s} {100 * v[i] / tot:5.1f} %   {v[i] / max(v[8], 1):8.0f} cycles per warp-step")
print(f"  total {tot / max(v[8], 1):.0f} cycles per warp-step")
sub = {6: "keys + bucket counts", 7: "scan", 12: "scatter + bucket sort", 13: "cell-sorted records", 14: "candidate lists"}
for i, nm in sub.items():
    print(f"  rebuild / {nm:24s} {100 * v[i] / max(v[0], 1):5.1f} % of the rebuild   {v[i] / max(v[9], 1):8.0f} cycles per warp-rebuild")
