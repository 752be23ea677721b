"""BASELINE configs[3]-like scene (many bodies, voxel-voxel contact on): per-step time of the automatic path as the scene grows."""
import sys, time
import numpy as np
import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W
for k, cube in ((3, 5), (5, 8), (8, 8), (12, 8)):
    model = gv.Model.from_vxa(W.contact_pile_vxa(k=k, cube=cube))
    deg = np.bincount(np.concatenate([model.array("link_neg"), model.array("link_pos")]), minlength=model.num_voxels)
    ns = int((deg < 6).sum())
    ctx = gv.Context(0, gv.FP32)
    ctx.upload([model])
    ctx.run(50)
    steps = 400
    ctx.run(steps)
    ms, launches = ctx.timing()
    r = ctx.results()[0]
    print(f"pile k={k} cube={cube}: {model.num_voxels} voxels, {ns} surface, {1e3 * ms / steps:8.1f} us/step, {model.num_voxels * steps / (ms * 1e-3):.3e} voxel-steps/s, "
          f"all-pairs {ns * (ns - 1) / 2 * steps / (ms * 1e-3):.3e} pair tests/s equivalent, hits {r.collision_count}, diverged {r.diverged}, launches {launches}", flush=True)
    ctx.close()
