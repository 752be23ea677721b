# One ncu metric pass over EVERY kernel of the engine (BASELINE north star: "every kernel ships with an ncu capture showing
# achieved HBM GB/s against B200 peak").  Short runs of each workload; per-launch CSVs -> gpurun_out/allk_<workload>_<tag>.csv,
# summarised by scripts/summarize_allkernels.py into profiles/<tag>_all_kernels.md.   Usage: bash scripts/gpu_allkernels.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
run() { # name, max launches, command...
  name=$1; cap=$2; shift 2
  timeout 600 ncu --metrics $M --clock-control none -k regex:^k_ -c $cap --csv --log-file gpurun_out/allk_${name}_$tag.csv "$@" > gpurun_out/allk_${name}_$tag.log 2>&1
  echo "$name rc=$? rows=$(wc -l < gpurun_out/allk_${name}_$tag.csv)"
}
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
run xenobot 40 $B --sim-steps 200
run cube_brick 60 $B --workload cube --sim-steps 20
run cube_stream 60 $B --workload cube --path 1 --sim-steps 8
run cube_fused 60 $B --workload cube --path 3 --sim-steps 8
run pile 400 $B --workload pile --sim-steps 30
run slab 200 python -m pytest tests/test_slab.py -q -x -m gpu -k "halo_exchange and 2"
