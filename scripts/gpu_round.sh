# One GPU session: parity tests, ad-hoc sweep, bench, ncu launch list + full capture.  Usage: bash scripts/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/pytest_$tag.log
timeout 300 python bench.py > gpurun_out/bench_$tag.log 2>&1; echo bench rc=$?; tail -2 gpurun_out/bench_$tag.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --sim-steps 100 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1; echo launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resident -c 2 -o gpurun_out/prof_resident_$tag python bench.py --sim-steps 200 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1; echo full rc=$?; tail -3 gpurun_out/ncu_full_$tag.log
timeout 300 python tests/gpu_check.py > gpurun_out/gpu_check_$tag.log 2>&1; tail -30 gpurun_out/gpu_check_$tag.log
