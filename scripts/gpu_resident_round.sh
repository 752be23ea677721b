# one measurement round for k_resident work: parity suite, phase profile (RES_PROFILE build), bench (+ skin sweep), ncu phases
tag=${1:-x}
python -m pytest tests -m gpu -q --timeout=1200 -x -k "not test_report_of_the_reference_generation" 2>&1 | tail -6 > gpurun_out/${tag}_tests.log
VX3B_LIB=gpuvoxels_b200/libvx3b200_prof.so python scripts/gpu_resident_phases.py > gpurun_out/${tag}_phases.log 2>&1
python bench.py --workloads none --no-cpu-baseline > gpurun_out/${tag}_bench.json 2>gpurun_out/${tag}_bench.err
for s in ${SKINS:-0.3}; do python bench.py --workloads none --no-cpu-baseline --opt contact_skin=$s --steps 3 > gpurun_out/${tag}_bench_skin$s.json 2>&1; done
[ -n "$NCU" ] && bash scripts/gpu_prof_phases.sh $tag >/dev/null 2>&1
tail -3 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_phases.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["value"], j["ms_per_step"], j["contact_rebuilds_per_world_rank0"])
    except Exception as e: print(f, "failed", e)
PY
