# Final numbers of a round without profiler passes: parity suite, smoke, every bench workload + the reference arm.  Usage: bash scripts/gpu_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_xeno_$tag.log 2>&1; echo bench xenobot rc=$?; tail -1 gpurun_out/bench_xeno_$tag.log | cut -c1-160
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.log 2>&1; echo bench reference rc=$?; tail -1 gpurun_out/bench_ref_$tag.log | cut -c1-160
timeout 600 python bench.py --workload cube --cube 100 --sim-steps 1000 --steps 3 --warmup 3 > gpurun_out/bench_cube_$tag.log 2>&1; echo bench cube rc=$?; tail -1 gpurun_out/bench_cube_$tag.log | cut -c1-160
timeout 600 python bench.py --workload pile --sim-steps 500 --steps 3 --warmup 3 > gpurun_out/bench_pile_$tag.log 2>&1; echo bench pile rc=$?; tail -1 gpurun_out/bench_pile_$tag.log | cut -c1-160
