# ncu --set full capture of k_resident (RES_MARKERS build) on 1,024 xenobot worlds x 200 steps; read here with scripts/ncu_phases.py
mkdir -p gpurun_out
tag=${1:-x}
VX3B_LIB=gpuvoxels_b200/libvx3b200_pm.so timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_resident -c 2 -o gpurun_out/prof_phases_$tag python bench.py --sim-steps 200 --steps 1 --warmup 0 --no-cpu-baseline --workloads none > gpurun_out/ncu_phases_$tag.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_phases_$tag.log | cut -c1-200; ls -la gpurun_out/prof_phases_$tag.ncu-rep
