# GPU session for the brick path: its parity tests first, then the cube bench on the fused (3) and automatic/brick (0) paths.
# Usage: bash scripts/gpu_round3.sh <tag> [pytest -k expression]
tag=${1:-r01f}; kexpr=${2:-"brick or million or BRICK or 5-"}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "$kexpr" > gpurun_out/pytest_$tag.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest_$tag.log
for p in 3 0; do
VX3B_DEBUG=1 timeout 400 python bench.py --workload cube --cube 100 --sim-steps 200 --steps 3 --warmup 3 --path $p --no-cpu-baseline > gpurun_out/bench_cube_p${p}_$tag.log 2>&1; echo cube path $p rc=$?; grep "vx3b\]" gpurun_out/bench_cube_p${p}_$tag.log | tail -3; tail -1 gpurun_out/bench_cube_p${p}_$tag.log | cut -c1-900
done
