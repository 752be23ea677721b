"""Host model (VXA/VXD reader, link construction, material constants, formulas) against the reference's own parser
output, frozen in tests/golden/*.vxbin (and re-checked live when oracle/_ref is present)."""
import math
import os

import numpy as np
import pytest

import gpuvoxels_b200 as gv
import oracle_lib as O
from gpuvoxels_b200 import worlds as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

VM = ['E', 'nu', 'rho', 'alphaCTE', 'muStatic', 'muKinetic', 'zetaInternal', 'zetaGlobal', 'zetaCollision', 'nomSize', 'gravMult', 'mass',
      'massInverse', 'sqrtMass', 'momentInertia', 'momentInertiaInverse', '_2xSqMxExS', '_2xSqIxExSxSxS', 'Cilia', 'fixed', 'isTarget',
      'isMeasured', 'matid', 'epsilonFail']
LM = ['E', 'nu', 'a1', 'a2', 'b1', 'b2', 'b3', 'sqA1', 'sqA2xIp', 'sqB1', 'sqB2xFMp', 'sqB3xIp', 'epsilonFail']


def assert_model_matches(m, md):
    assert np.array_equal(m.array('lattice'), md['voxels'][:, :3])
    assert np.array_equal(m.array('voxel_mat'), md['voxels'][:, 3])
    assert np.array_equal(np.stack([m.array('link_neg'), m.array('link_pos'), m.array('link_axis')], 1), md['links'])
    assert np.array_equal(m.array('phase_offset'), md['voxel_extra'][:, 0])
    assert np.array_equal(m.array('base_cilia'), md['voxel_extra'][:, 1:4])
    for i, mat in enumerate(m.voxel_materials()):
        mine = np.array([getattr(mat, n) for n in VM], dtype=np.float64)
        assert np.array_equal(mine, md['voxel_mats'][i]), (i, mine, md['voxel_mats'][i])  # bit-exact float constants
    lm, lmi = m.link_materials(), m.array('link_mat')
    for l in range(m.num_links):
        mine = np.array([getattr(lm[lmi[l]], n) for n in LM], dtype=np.float64)
        assert np.array_equal(mine, md['link_mats'][l, :13]), l
    env = md['env'][0]
    d = m.desc
    assert d.vox_size == env[0] and d.dt_frac == env[1] and bool(d.floor_enabled) == bool(env[5])
    assert bool(d.vary_temp_enabled) == bool(env[8]) and d.temp_amplitude == env[10] and d.temp_period == env[11]


def golden_cases():
    from golden.make_golden import CASES
    return CASES


@pytest.mark.parametrize("name", ["example_vx2_1000", "small_vx3_2000", "xeno3_vx3_cilia_800", "cube8_vx2_400"])
def test_model_matches_reference_golden(name):
    vxa, vxd, *_ = golden_cases()[name]
    g = O.read_vxbin(os.path.join(GOLD, name + ".vxbin"))
    md = {k[len("model_"):]: v for k, v in g.items() if k.startswith("model_")}
    assert_model_matches(gv.Model.from_vxa(vxa, vxd), md)


def test_example_vxa_known_answers(example_vxa):
    # 140 voxels, 316 links, 10 materials, voxSize 0.01, DtFrac 0.3, gravity multiplier 1.00034
    m = gv.Model.from_vxa(example_vxa)
    assert (m.num_voxels, m.num_links, m.desc.num_voxel_mats) == (140, 316, 10)
    assert m.desc.vox_size == 0.01 and m.desc.dt_frac == 0.3
    assert abs(m.voxel_materials()[0].gravMult - 1.00034) < 1e-5
    w = O.OracleWorld(m)
    w.run(1)
    assert abs(w.result().dt / 0.3 - 0.000290576) < 1e-9  # recommendedTimeStep()


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("maker", [lambda: (W.small_block_vxa(), None), lambda: (W.xenobot_base_vxa(), W.xenobot_vxd(11)),
                                   lambda: (W.contact_pile_vxa(), None), lambda: (W.cube_vxa(5), None)])
def test_model_matches_reference_live(maker):
    vxa, vxd = maker()
    m = gv.Model.from_vxa(vxa, vxd)
    _, md, _ = O.run_reference(m.merged_vxa(), 0, want_model=True)
    assert_model_matches(m, md)


REF_INPUTS = os.path.join(ROOT, "tests", "data", "ref_inputs")


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case", [("e.circle_push_data901", f"bot_{i:04d}.vxd") for i in range(11)] + [("d.memleaking_data1", "bot_0000.vxd")],
                         ids=lambda c: c[0][0] + "_" + c[1][:8])
def test_model_of_the_reference_experiment_inputs_matches_the_reference_parser(case):
    """the reference's own checked-in experiment inputs (Experiments/xenobot/e.circle_push/data901, d.memleaking/data1): every
    voxel, link, per-voxel phase / cilia vector and every float material constant is the one the reference's reader builds"""
    d, f = case
    m = gv.Model.from_files(os.path.join(REF_INPUTS, d, "base.vxa"), os.path.join(REF_INPUTS, d, f))
    _, md, _ = O.run_reference(m.merged_vxa(), 0, want_model=True)
    assert_model_matches(m, md)
    assert m.num_voxels >= 34


def test_vxd_replaces_subtree_and_keeps_the_rest():
    base = W.xenobot_base_vxa()
    m0 = gv.Model.from_vxa(base)
    m1 = gv.Model.from_vxa(base, W.xenobot_vxd(5))
    assert m0.num_voxels == 4 and m1.num_voxels > 50
    assert m1.desc.enable_cilia == 1 and m1.desc.dt_frac == 1.0
    assert np.abs(m1.array("base_cilia")).max() > 0
    # VXD may also replace a scalar deep in the tree
    vxd = '<VXD><DtFrac replace="VXA.Simulator.Integration.DtFrac">0.25</DtFrac></VXD>'
    assert gv.Model.from_vxa(base, vxd).desc.dt_frac == 0.25


def test_formula_tokens_follow_the_reference_order():
    # <mtSUB><mtVAR>t</mtVAR><mtCONST>5</mtCONST></mtSUB> -> reversed level order: CONST 5, VAR t, SUB, END
    m = gv.Model.from_vxa(W.xenobot_base_vxa(stop_seconds=5.0))
    t = m.desc.stop_condition
    assert [t[i].op for i in range(4)] == [1, 4, 6, 0]
    assert t[0].value == 5.0 and t[1].value == 4.0
    f = m.desc.fitness
    n = next(i for i in range(64) if f[i].op == 0)
    assert n == 16
    # run the reference's FIFO queue machine (Utils/VX3_MathTree.h:50-193) over our tokens: |angle| + ln(hit+1) + ln(x^2+y^2)
    var = {0: 3.0, 1: 4.0, 3: 2.0, 5: -0.5}
    values, pc = [], 0
    for i in range(n):
        op, val = f[i].op, f[i].value
        if op == 1:
            values.append(val)
        elif op == 4:
            values.append(var[int(val)])
        elif op in (5, 7):
            a, b = values[pc + 1], values[pc]
            pc += 2
            values.append(a + b if op == 5 else a * b)
        elif op == 15:
            values.append(math.log(values[pc]))
            pc += 1
        elif op == 17:
            values.append(abs(values[pc]))
            pc += 1
        else:
            raise AssertionError(op)
    assert abs(values[pc] - (0.5 + math.log(3.0) + math.log(25.0))) < 1e-12


def test_bad_inputs_fail_with_a_message():
    with pytest.raises(gv.VX3Error):
        gv.Model.from_vxa("<VXA><Simulator>")
    with pytest.raises(gv.VX3Error):
        gv.Model.from_vxa("<NotVXA/>")
    bad = W.small_block_vxa().replace("<Elastic_Mod>5000000.0</Elastic_Mod>", "<Elastic_Mod>0</Elastic_Mod>")
    with pytest.raises(gv.VX3Error):
        gv.Model.from_vxa(bad)
    m = gv.Model.from_vxa(W.small_block_vxa().replace("<Simulator>", "<Simulator><AttachDetach><EnableAttach>1</EnableAttach></AttachDetach>", 1))
    assert m.desc.enable_attach == 1  # (refused until round 2: SURVEY 8f-3)


def test_builder_link_order_matches_cvoxelyze_add_order():
    # voxels added out of lattice order: links appear when the second end arrives, directions X_POS..Z_NEG
    m = gv.Model.new(0.001)
    mat = m.add_material(1e6, 1e3)
    for p in [(1, 0, 0), (0, 0, 0), (1, 1, 0), (0, 1, 0)]:
        m.set_voxel(mat, *p)
    m.finalize()
    ends = list(zip(m.array("link_neg"), m.array("link_pos"), m.array("link_axis")))
    assert ends == [(1, 0, 0), (0, 2, 1), (3, 2, 0), (1, 3, 1)]
