# Why is the per-step time in bench.py above the ncu kernel duration?  Long run with fast clock sampling.
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader -lms 100 > gpurun_out/clk_probe.csv &
SMI=$!
VX3B_DEBUG=1 timeout 600 python bench.py --workload cube --cube 100 --sim-steps 2000 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cube_long.log 2>&1
kill $SMI
tail -1 gpurun_out/bench_cube_long.log | cut -c1-200
sort gpurun_out/clk_probe.csv | uniq -c | sort -rn | head -12
