# ncu capture of k_brick late in a run (the body has hit the floor and is ringing), to compare with the early-step capture.
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_brick$ -s 700 -c 1 -o gpurun_out/prof_brick_late_$tag python bench.py --workload cube --cube 100 --sim-steps 720 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_brick_late_$tag.log 2>&1; echo full rc=$?; tail -2 gpurun_out/ncu_full_brick_late_$tag.log | cut -c1-200
