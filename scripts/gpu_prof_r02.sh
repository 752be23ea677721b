# ncu evidence of a round: launch lists (shares of a step) of the three single-GPU workloads and one full-set capture each of
# k_resident and k_brick (and k_fused on the pile), summarised into profiles/ by scripts/summarize_profiles.py <tag>.  Usage: bash scripts/gpu_prof_r02.sh <tag>
tag=${1:-r02u}
mkdir -p gpurun_out
N="--no-cpu-baseline --workloads none"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_xeno_$tag.csv python bench.py --sim-steps 100 --steps 2 --warmup 1 $N > gpurun_out/ncu_launch_xeno_$tag.log 2>&1; echo launches xeno rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cube_$tag.csv python bench.py --workload cube --cube 100 --sim-steps 20 --steps 2 --warmup 1 $N > gpurun_out/ncu_launch_cube_$tag.log 2>&1; echo launches cube rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_pile_$tag.csv python bench.py --workload pile --sim-steps 60 --steps 1 --warmup 0 $N > gpurun_out/ncu_launch_pile_$tag.log 2>&1; echo launches pile rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resident -c 2 -o gpurun_out/prof_resident_$tag python bench.py --sim-steps 200 --steps 1 --warmup 0 $N > gpurun_out/ncu_full_resident_$tag.log 2>&1; echo full resident rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_brick$ -s 5 -c 1 -o gpurun_out/prof_brick_$tag python bench.py --workload cube --cube 100 --sim-steps 20 --steps 1 --warmup 0 $N > gpurun_out/ncu_full_brick_$tag.log 2>&1; echo full brick rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_fused$ -s 20 -c 1 -o gpurun_out/prof_fused_$tag python bench.py --workload pile --sim-steps 40 --steps 1 --warmup 0 $N > gpurun_out/ncu_full_fused_$tag.log 2>&1; echo full fused rc=$?
ls -la gpurun_out/*_$tag*
