"""Where does a slab's time step go?  ONE slab of a body (with its ghost layers) on one GPU, no neighbours: the step kernel on the
slab's shape alone; the difference to the multi-GPU figure of bench.py --workload slab is the exchange plus the wait for the
slowest neighbour.
usage: python scripts/gpu_slab_probe.py EDGE N_PARTS PART [STEPS]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gpuvoxels_b200 as gv  # noqa: E402
from gpuvoxels_b200 import slab as S  # noqa: E402
from gpuvoxels_b200 import worlds as W  # noqa: E402

edge, n_parts, part_id = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 64
vxa = W.cube_vxa(edge)
part = S.native_part(vxa, None, part_id, n_parts)
print(f"slab {part_id}/{n_parts} of {edge}^3: {part.num_own} own + {part.model.num_voxels - part.num_own} ghosts, {part.model.num_links} links")


def timed(loop):
    run = S.SlabRunner(part, 0, gv.FP32, part.body_dt)
    if loop:
        ep = run.open_endpoints()
        run.connect(ep["hi"], ep["lo"], same_process=True)
    run.run(steps)  # warm-up, graph capture
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(run.stream):
        ev[0].record()
    run.run(steps)
    with torch.cuda.stream(run.stream):
        ev[1].record()
    torch.cuda.synchronize()
    us = 1e3 * ev[0].elapsed_time(ev[1]) / steps
    print(f"  {'step + looped-back push/pull' if loop else 'step kernel alone':32s} {us:8.1f} us per time step  ({part.num_own / us * 1e6:.3e} own voxel-steps/s), timeouts {run.ctx.halo_timeouts()}")
    run.close()


timed(False)
