"""signals, attachment, detachment, removal.  The oracle restates them in the serial thread order of the reference's own
Voxelyze3 step code built for the host (oracle/_ref/vx3_ref) - ONE valid execution of kernels that race on a GPU - and must then
equal it bit for bit: positions, every signal scalar of every voxel, the grown link list, the removed / detached flags."""
import os

import numpy as np
import pytest

import gpuvoxels_b200 as gv
import oracle_lib as O
from gpuvoxels_b200 import worlds as W
from golden.make_golden_vx3 import CASES, merged

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INPUTS = os.path.join(ROOT, "tests", "data", "ref_inputs")


def test_the_reference_experiment_inputs_that_use_these_features_load():
    """Experiments/xenobot/c.signaling/data1|data2 and h.target_only/data1 (SURVEY's config-3 exemplar) were refused in round 1"""
    for d, f, want in (("c.signaling_data1", "rob.vxd", dict(enable_signals=1, enable_attach=0)),
                       ("c.signaling_data2", "rob.vxd", dict(enable_signals=1, enable_cilia=0)),
                       ("h.target_only_data1", "bot_0000.vxd", dict(enable_signals=1, enable_attach=1, enable_detach=1, safety_guard=100000))):
        m = gv.Model.from_files(os.path.join(REF_INPUTS, d, "base.vxa"), os.path.join(REF_INPUTS, d, f))
        for k, v in want.items():
            assert getattr(m.desc, k) == v, (d, k)
    m = gv.Model.from_files(os.path.join(REF_INPUTS, "h.target_only_data1", "base.vxa"), os.path.join(REF_INPUTS, "h.target_only_data1", "bot_0000.vxd"))
    assert m.num_voxels == 2964
    vm = m.voxel_materials()
    sticky = [i for i, v in enumerate(vm) if v.sticky]
    assert sticky and all(any((l.mat1, l.mat2) == (i, i) for l in m.link_materials()) for i in sticky)  # the material attachment will need
    assert abs(vm[1].inactivePeriod - 0.5) < 1e-12 and abs(vm[1].signalTimeDelay - 0.01) < 1e-12  # the cilia material of the bodies
    assert m.desc.attach_condition[0][0].op == 1 and m.desc.attach_condition[1][0].op == 0  # Condition_0 = CONST 1, the others absent


@pytest.mark.skipif(not O.have_reference_vx3(), reason="oracle/_ref/vx3_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name,steps", [("vx3ref_signaling_data1_3000", 2500), ("vx3ref_signaling_data2_3000", 2500), ("vx3ref_sticky_pebbles_3000", 2000),
                                        ("vx3ref_breaking_chain_3000", 1500), ("vx3ref_target_only_300", 120)])
def test_oracle_is_bit_identical_to_the_live_vx3_reference(name, steps):
    vxa, vxd, _, _, _ = CASES[name]
    ref, _ = O.run_reference_vx3(merged(vxa, vxd), steps, dump_at=(1, steps // 2))
    w = O.OracleWorld(gv.Model.from_vxa(vxa, vxd))
    done = 0
    for c, sfx in ((1, "_1"), (steps // 2, f"_{steps // 2}"), (steps, "")):
        w.run(c - done)
        done = c
        s = w.state()
        for k in ("pos", "orient", "linMom", "angMom"):
            assert np.array_equal(s[k], ref[k + sfx]), (k, sfx)
        sig, fl = w.signals()
        assert np.array_equal(sig, ref["signals" + sfx]), sfx
        assert np.array_equal(w.topology(), ref["links" + sfx]), sfx
        assert np.array_equal(fl[:, 0], ref["vflags" + sfx][:, 1]) and np.array_equal(fl[:, 1], ref["vflags" + sfx][:, 2])
        assert w.result().collision_count == ref["scalars" + sfx][0, 4]


def test_the_synthetic_worlds_do_exercise_the_features():
    g = O.read_vxbin(os.path.join(ROOT, "tests", "golden", "vx3ref_sticky_pebbles_3000.vxbin"))
    assert len(g["links_200"]) == 5 and len(g["links"]) == 6       # four raft links, then the two pebbles attach
    assert g["links"][4:, 5].max() == 0                             # isNewLink has counted down from SafetyGuard to 0
    assert sorted(g["links"][4:, 2]) == [2, 2]                      # both new links run along z: the pebbles sit on the raft
    assert g["signals"][:, 0].max() > 50                            # the pace maker's signal is travelling
    g = O.read_vxbin(os.path.join(ROOT, "tests", "golden", "vx3ref_breaking_chain_3000.vxbin"))
    assert g["links"][:, 4].sum() >= 1 and g["vflags"][:, 1].sum() == 2 and g["links"][:, 3].sum() >= 2  # detached, removed voxels, removed links
    g = O.read_vxbin(os.path.join(ROOT, "tests", "golden", "vx3ref_signaling_data1_3000.vxbin"))
    assert g["scalars"][0, 4] > 1000 and g["signals"][:, 0].max() == 100  # target hits raise signals
