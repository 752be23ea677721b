# ncu --set full capture of the resident kernel on the bench workload.  Usage: bash scripts/gpu_prof.sh <tag> [extra bench args]
tag=${1:-x}; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resident -c 2 -o gpurun_out/prof_resident_$tag python bench.py --sim-steps 200 --steps 1 --warmup 0 --no-cpu-baseline "$@" > gpurun_out/ncu_full_$tag.log 2>&1; echo full rc=$?; tail -2 gpurun_out/ncu_full_$tag.log | cut -c1-300
