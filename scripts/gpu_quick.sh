# Quick GPU check of a kernel change: parity suite + default bench line.  Usage: bash scripts/gpu_quick.sh <tag>
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_$tag.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/bench_$tag.log | cut -c1-400
