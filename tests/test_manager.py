"""gv.SimulationManager (mirror of VX3_SimulationManager, VX3/VX3_SimulationManager.cuh:15-46): host-side bookkeeping over
unequal sub-batches, with the device context replaced by a stand-in (the GPU run itself is covered by the -m gpu suite)."""
import math
import os
import tempfile

import numpy as np

import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W


class _Result:
    def __init__(self, fitness):
        self.fitness_score = fitness


class _FakeContext:
    """results() and read_state(k) answer with the voxel count of the k-th uploaded model, so a mix-up of worlds shows"""
    created = []

    def __init__(self, device, precision):
        self.device = device
        _FakeContext.created.append(device)

    def upload(self, models):
        self.models = models

    def run(self, steps, use_stop_formula=False):
        assert use_stop_formula and steps == 1000000

    def results(self):
        return [_Result(float("nan") if m.num_voxels % 5 == 0 else float(m.num_voxels)) for m in self.models]

    def read_state(self, k):
        return {"pos": np.full((self.models[k].num_voxels, 3), float(self.models[k].num_voxels))}

    def close(self):
        pass


def test_unequal_sub_batches_keep_every_world_with_its_own_positions(monkeypatch):
    monkeypatch.setattr(gv, "Context", _FakeContext)
    _FakeContext.created = []
    with tempfile.TemporaryDirectory() as td:
        base = os.path.join(td, "base.vxa")
        open(base, "w").write(W.xenobot_base_vxa().replace("<Simulator>", "<Simulator><SavePositionOfAllVoxels>1</SavePositionOfAllVoxels>", 1))
        files = []
        for i in range(5):
            files.append(f"bot_{i:04d}.vxd")
            open(os.path.join(td, files[-1]), "w").write(W.xenobot_vxd(60 + i))
        sm = gv.SimulationManager([files[:3], files[3:]], base, td, 2)  # 3 + 2: the case the per-device index used to get wrong
        out = sm.start()
    assert sorted(_FakeContext.created) == [0, 1]
    assert sorted(e["vxa_filename"] for e in out) == files
    for e in out:
        nv = e["model"].num_voxels
        assert e["pos"].shape == (nv, 3) and np.all(e["pos"] == nv)  # the positions of THIS world
    fits = [e["result"].fitness_score for e in out]
    finite = [f for f in fits if not math.isnan(f)]
    assert finite == sorted(finite, reverse=True) and all(math.isnan(f) for f in fits[len(finite):])  # descending, NaN last
