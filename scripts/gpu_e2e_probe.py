"""Where the end-to-end time of one generation goes: upload / run / results, wall clock, through the Python shim."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
base = W.xenobot_base_vxa()
models = [gv.Model.from_vxa(base, W.xenobot_vxd(i)) for i in range(n)]
ctx = gv.Context(0, gv.FP32)
ctx.upload(models); ctx.run(1000)
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ctx.upload(models)
    t1 = time.perf_counter(); ctx.run(1000)
    t2 = time.perf_counter(); res = ctx.results()
    t3 = time.perf_counter(); fit = torch.tensor([r.fitness_score for r in res], dtype=torch.float64, device="cuda"); _ = fit.cpu()
    t4 = time.perf_counter()
    ms, launches = ctx.timing()
    print(f"upload {1e3*(t1-t0):6.1f} ms  run {1e3*(t2-t1):6.1f} ms (device {ms:6.1f})  results {1e3*(t3-t2):5.1f} ms  fitness->cuda->cpu {1e3*(t4-t3):5.1f} ms  total {1e3*(t4-t0):6.1f}", flush=True)
