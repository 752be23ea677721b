# compute-sanitizer over the kernels round 2 added or changed: memcheck on the slab exchanges (fused into k_brick, staged push / pull),
# the contact grid (epoch heads, list build + narrow phase in one kernel, overflow path), the per-body brick cut and the pipelined
# generation runner; racecheck on k_brick (bulk copies a brick ahead) and k_fused.  Usage: bash scripts/gpu_sanitize_r02.sh
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, "tests")
import numpy as np
import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W
import test_slab as TS
what = sys.argv[1]
if what == "slab":
    for fused in (True, False):
        got = TS.run_slabs_device_exchange(W.cube_vxa(9), None, 3, 40, gv.FP32, dict(brick_x=4, brick_y=3, brick_z=2), (40,), fused)
        print("slab fused" if fused else "slab staged", float(np.abs(got["pos"]).max()))
elif what == "contact":
    models = [gv.Model.from_vxa(W.contact_pile_vxa(k=3, cube=4)), gv.Model.from_vxa(W.contact_pile_vxa(k=2, cube=3, seed=3))]
    for path in (1, 3):
        for skin, cap in ((0.15, 0), (0.5, 1)):
            ctx = gv.Context(0, gv.FP32)
            ctx.set_option("path", path); ctx.set_option("contact_skin", skin); ctx.set_option("contact_list_cap", cap)
            ctx.upload(models)
            ctx.run(120)
            print("contact path", path, skin, cap, [(r.steps, r.contact_rebuilds, r.contact_overflows, r.collision_count) for r in ctx.results()])
            ctx.close()
elif what == "gen":
    models = [gv.Model.from_vxa(W.xenobot_base_vxa(stop_seconds=0.01), W.xenobot_vxd(40 + i)) for i in range(9)]
    gen = gv.Generation(0, gv.FP32, 3)
    print("gen", [r.steps for r in gen.run(models, 1000000, True)])
    gen.close()
elif what == "race":
    models = [gv.Model.from_vxa(W.cube_vxa(9)), gv.Model.from_vxa(W.contact_pile_vxa(k=2, cube=3, seed=3))]
    for path in (5, 3):
        ctx = gv.Context(0, gv.FP32)
        ctx.set_option("path", path)
        ctx.set_option("brick_x", 4); ctx.set_option("brick_y", 4); ctx.set_option("brick_z", 2)
        ctx.upload(models)
        ctx.run(10)
        print("race path", path, [r.steps for r in ctx.results()])
        ctx.close()
PY
for w in slab contact gen; do
PYTHONPATH=. timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python /tmp/san.py $w > gpurun_out/memcheck_$w.log 2>&1
echo "memcheck $w: $(grep 'ERROR SUMMARY' gpurun_out/memcheck_$w.log | tail -1) | $(grep -v '^=' gpurun_out/memcheck_$w.log | tail -3 | tr '\n' ' ' | cut -c1-300)"
done
PYTHONPATH=. timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python /tmp/san.py race > gpurun_out/racecheck_r02.log 2>&1
echo "racecheck: $(grep -c 'Race reported' gpurun_out/racecheck_r02.log) race reports; $(grep 'RACECHECK SUMMARY' gpurun_out/racecheck_r02.log | tail -1) | $(grep -v '^=' gpurun_out/racecheck_r02.log | tail -2 | tr '\n' ' ')"
grep -m 6 -A1 "Race reported" gpurun_out/racecheck_r02.log | cut -c1-260
