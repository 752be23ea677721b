# Sweep of the contact-list skin on the xenobot generation (bench.py --opt contact_skin=...).  Usage: bash scripts/gpu_skin_sweep.sh
mkdir -p gpurun_out
for s in 0.15 0.3 0.45 0.6 0.9; do
  echo "== skin $s" >> gpurun_out/skin_sweep.log
  timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt contact_skin=$s 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('skin', '$s', 'value %.3e' % j['value'], 'ms', round(j['ms_per_step'], 1), 'rebuilds/world', j.get('contact_rebuilds_per_world_rank0'), 'overflows', j.get('contact_overflows_rank0'), 'e2e %.3e' % j['e2e']['value'])" >> gpurun_out/skin_sweep.log 2>&1
done
cat gpurun_out/skin_sweep.log
