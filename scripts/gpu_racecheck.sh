# compute-sanitizer racecheck (shared-memory hazards) of the three step kernels on small configurations.
mkdir -p gpurun_out
cat > /tmp/rc.py <<'PY'
import sys
import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W
path = int(sys.argv[1])
models = [gv.Model.from_vxa(W.cube_vxa(9)), gv.Model.from_vxa(W.xenobot_base_vxa(enable_collision=(path == 3)), W.xenobot_vxd(9)), gv.Model.from_vxa(W.small_block_vxa())]
ctx = gv.Context(0, gv.FP32)
ctx.set_option("path", path)
ctx.set_option("brick_x", 4); ctx.set_option("brick_y", 4); ctx.set_option("brick_z", 2)
ctx.upload(models)
ctx.run(12)
print("ran", [r.steps for r in ctx.results()])
PY
for p in 5 3 2; do
PYTHONPATH=. timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python /tmp/rc.py $p > gpurun_out/race_p$p.log 2>&1
echo "path $p: $(grep -c 'Race reported' gpurun_out/race_p$p.log) race reports; $(grep 'RACECHECK SUMMARY' gpurun_out/race_p$p.log)"; grep -m 6 -A1 "Race reported" gpurun_out/race_p$p.log | cut -c1-260
done
