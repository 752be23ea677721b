# GPU session for the fused tile path: parity tests, both bench workloads, launch list + full capture of k_fused.  Usage: bash scripts/gpu_round2.sh <tag>
tag=${1:-r01e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_$tag.log
timeout 400 python bench.py > gpurun_out/bench_xeno_$tag.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/bench_xeno_$tag.log | cut -c1-600
for p in 0 1; do
VX3B_DEBUG=1 timeout 400 python bench.py --workload cube --cube 100 --sim-steps 200 --steps 3 --warmup 3 --path $p --no-cpu-baseline > gpurun_out/bench_cube_p${p}_$tag.log 2>&1; echo cube path $p rc=$?; tail -1 gpurun_out/bench_cube_p${p}_$tag.log | cut -c1-700
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cube_$tag.csv python bench.py --workload cube --cube 100 --sim-steps 20 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_cube_$tag.log 2>&1; echo launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 5 -c 2 -o gpurun_out/prof_fused_$tag python bench.py --workload cube --cube 100 --sim-steps 20 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_fused_$tag.log 2>&1; echo full rc=$?; tail -2 gpurun_out/ncu_full_fused_$tag.log | cut -c1-300
