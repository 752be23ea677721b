# ncu capture of k_brick on the 100^3 cube: launch list + one full-set capture with source counters.  Usage: bash scripts/gpu_prof_brick.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_brick$ -s 5 -c 1 -o gpurun_out/prof_brick_$tag python bench.py --workload cube --cube 100 --sim-steps 20 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_brick_$tag.log 2>&1; echo full rc=$?; tail -2 gpurun_out/ncu_full_brick_$tag.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cube_$tag.csv python bench.py --workload cube --cube 100 --sim-steps 20 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_cube_$tag.log 2>&1; echo launches rc=$?
