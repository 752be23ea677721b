# Full GPU round: parity suite, every bench workload (+ reference arm), launch lists and full captures of the dominant kernels.
# Usage: bash scripts/gpu_round5.sh <tag>
tag=${1:-r01r}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_$tag.log
timeout 600 python bench.py > gpurun_out/bench_xeno_$tag.log 2>&1; echo bench xenobot rc=$?; tail -1 gpurun_out/bench_xeno_$tag.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.log 2>&1; echo bench reference rc=$?; tail -1 gpurun_out/bench_ref_$tag.log | cut -c1-200
VX3B_DEBUG=1 timeout 600 python bench.py --workload cube --cube 100 --sim-steps 1000 --steps 3 --warmup 3 > gpurun_out/bench_cube_$tag.log 2>&1; echo bench cube rc=$?; tail -1 gpurun_out/bench_cube_$tag.log | cut -c1-200
timeout 600 python bench.py --workload pile --sim-steps 500 --steps 3 --warmup 3 > gpurun_out/bench_pile_$tag.log 2>&1; echo bench pile rc=$?; tail -1 gpurun_out/bench_pile_$tag.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_xeno_$tag.csv python bench.py --sim-steps 100 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_xeno_$tag.log 2>&1; echo launches xeno rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cube_$tag.csv python bench.py --workload cube --cube 100 --sim-steps 20 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_cube_$tag.log 2>&1; echo launches cube rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_pile_$tag.csv python bench.py --workload pile --sim-steps 20 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_pile_$tag.log 2>&1; echo launches pile rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_brick$ -s 700 -c 1 -o gpurun_out/prof_brick_$tag python bench.py --workload cube --cube 100 --sim-steps 720 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_brick_$tag.log 2>&1; echo full brick rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resident -c 6 -o gpurun_out/prof_resident_$tag python bench.py --sim-steps 200 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_resident_$tag.log 2>&1; echo full resident rc=$?
