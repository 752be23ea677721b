"""Pins the CPU oracle (oracle/vx3_oracle.cpp) against the unmodified reference CPU library: frozen outputs in
tests/golden (always) and live runs of oracle/_ref/vx2_ref (when it has been built from /root/reference)."""
import os

import numpy as np
import pytest

import gpuvoxels_b200 as gv
import oracle_lib as O
from gpuvoxels_b200 import worlds as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def prepared_model(vxa, vxd, mode, cilia):
    m = gv.Model.from_vxa(vxa, vxd)
    m.set_world_param("enable_collision", 0)  # the harness cannot emulate VX3 contact (oracle header: parity unpinned there)
    m.set_world_param("enable_cilia", 1 if cilia else 0)
    if mode == "vx2":
        m.set_world_param("vx2_compat", 1)
        m.set_world_param("vary_temp", 0)
    return m


def compare(w, ref, tol_pos, tol_other):
    s, l = w.state(), w.links()
    assert O.rel_err(s["pos"], ref["pos"]) < tol_pos
    assert O.rel_err(s["orient"], ref["orient"]) < tol_pos
    assert O.rel_err(s["linMom"], ref["linMom"]) < tol_other
    assert O.rel_err(s["angMom"], ref["angMom"]) < tol_other
    assert O.rel_err(l["pos2"], ref["link_pos2"]) < tol_other
    assert O.rel_err(l["angle2v"], ref["link_angle2v"]) < tol_other
    assert np.array_equal(l["flags"], ref["link_flags"][:, 0])


@pytest.mark.parametrize("name", ["example_vx2_1000", "example_vx3_1000", "small_vx3_2000", "xeno3_vx3_cilia_800", "cube8_vx2_400"])
def test_oracle_matches_reference_golden(name):
    from golden.make_golden import CASES
    vxa, vxd, steps, mode, cilia = CASES[name]
    g = O.read_vxbin(os.path.join(GOLD, name + ".vxbin"))
    ref = {k[len("state_"):]: v for k, v in g.items() if k.startswith("state_")}
    w = O.OracleWorld(prepared_model(vxa, vxd, mode, cilia))
    assert w.run(steps) == steps
    assert w.result().dt == ref["scalars"][0, 1]  # the float-truncated dt is bit-identical
    # the harness injects cilia through CVX_External's float force, so that case agrees to float precision only
    if name.startswith("cube"):
        # a cube dropped flat is a symmetric, laterally unstable configuration: last-bit differences pick the direction the
        # friction state machine breaks symmetry in (bit-exact for 2 steps, 1e-13 at step 3, then amplified) -> north-star tolerance
        compare(w, ref, 1e-5, 1e-2)
    else:
        compare(w, ref, 1e-5 if cilia else 1e-9, 1e-2 if cilia else 1e-6)


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
def test_one_step_is_bit_exact(example_vxa):
    m = prepared_model(example_vxa, None, "vx2", False)
    ref, _, _ = O.run_reference(m.merged_vxa(), 1)
    w = O.OracleWorld(m)
    w.run(1)
    s = w.state()
    for k in ("pos", "orient", "linMom", "angMom"):
        assert np.array_equal(s[k], ref[k]), k


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("mode", ["vx2", "vx3"])
def test_oracle_matches_reference_live(mode):
    vxa, vxd = W.xenobot_base_vxa(), W.xenobot_vxd(21)
    m = prepared_model(vxa, vxd, mode, False)
    ref, _, _ = O.run_reference(m.merged_vxa(), 600, vx3=(mode == "vx3"))
    w = O.OracleWorld(m)
    w.run(600)
    compare(w, ref, 1e-8, 1e-5)


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
def test_contact_force_matches_the_reference_collision_system_on_a_two_body_drop():
    """Pins the per-pair contact force (VX3_Collision.cu:3-31 = VX_Collision.cpp:42-56: envelope 0.625, series stiffness, mean
    damping, signs) against the UNMODIFIED reference: a voxel dropped from two cells up onto a voxel resting on the floor -
    one watched pair in VX3's all-pairs scheme and in Voxelyze2's watch lists alike - through impact, bounce and rest.
    Voxelyze2 forms the pair offset in float (VX_Collision.cpp:44), so agreement is at float, not double, precision.  What this
    does NOT pin is VX3's pair selection (all surface pairs, depth-1 exclusion), which Voxelyze2 does differently."""
    from golden.make_golden import two_body_model
    m = two_body_model()
    free_fall, _, _ = O.run_reference(m.merged_vxa(), 60)
    for steps in (40, 60, 100, 300, 1000):
        ref, _, _ = O.run_reference(m.merged_vxa(), steps, vx2_collisions=True)
        w = O.OracleWorld(m)
        w.run(steps)
        s = w.state()
        assert len(w.collision_pairs()) == 1
        assert np.abs(s["pos"] - ref["pos"]).max() < 1e-9, steps  # metres; the drop is 0.01 m, the contact overlap ~5e-4 m
        assert O.rel_err(s["pos"], ref["pos"]) < 1e-7
        if steps <= 100:
            assert O.rel_err(s["linMom"], ref["linMom"]) < 1e-5
    w = O.OracleWorld(m)
    w.run(60)
    assert np.abs(w.state()["pos"] - free_fall["pos"]).max() > 5e-3  # the contact is what holds the upper voxel: without it it falls through


def test_contact_force_matches_the_frozen_reference_run():
    """the same pin from the committed fixture (tests/golden/two_body_drop_vx2col.vxbin), for boxes without the reference binary"""
    from golden.make_golden import TWO_BODY_STEPS, two_body_model
    g = O.read_vxbin(os.path.join(GOLD, "two_body_drop_vx2col.vxbin"))
    m = two_body_model()
    for steps in TWO_BODY_STEPS:
        w = O.OracleWorld(m)
        w.run(steps)
        s = w.state()
        assert np.abs(s["pos"] - g[f"pos_{steps}"]).max() < 1e-9, steps
        if steps <= 100:
            assert O.rel_err(s["linMom"], g[f"linMom_{steps}"]) < 1e-5


def test_multi_pair_contact_matches_the_frozen_reference_run():
    """A 3x3x2 block landing on two pebbles, tipping and pushing them aside (tests/golden/block_on_pebbles_vx2col.vxbin: the
    unmodified reference with ITS collision system): several watched pairs at once, contact + floor friction, 2,000 steps."""
    from golden.make_golden import PEBBLE_STEPS, pebbles_model
    g = O.read_vxbin(os.path.join(GOLD, "block_on_pebbles_vx2col.vxbin"))
    m = pebbles_model()
    touched = 0
    for steps in PEBBLE_STEPS:
        w = O.OracleWorld(m)
        w.run(steps)
        s = w.state()
        touched += len(w.collision_pairs())
        assert np.abs(s["pos"] - g[f"pos_{steps}"]).max() < 5e-8, steps  # metres (voxel 0.01 m; Voxelyze2 forms pair offsets in float)
        assert np.abs(s["orient"] - g[f"orient_{steps}"]).max() < 1e-5, steps
    assert touched >= 4


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
def test_multi_pair_contact_matches_the_reference_live():
    from golden.make_golden import pebbles_model
    m = pebbles_model()
    ref, _, _ = O.run_reference(m.merged_vxa(), 700, vx2_collisions=True)
    w = O.OracleWorld(m)
    w.run(700)
    assert np.abs(w.state()["pos"] - ref["pos"]).max() < 5e-8


def test_vx3_cilia_body_pushing_a_target_matches_the_frozen_reference_run():
    """VX3 semantics (double dt, cilia, angMom zeroing) WITH contact against the reference (harness --vx3 + its own collision
    system; tests/golden/cilia_pusher_vx3_vx2col.vxbin): a cilia body walks into a target pebble and pushes it 1.4 m along the
    floor in 4,000 steps.  The harness injects cilia through CVX_External's float force, hence float-level agreement."""
    from golden.make_golden import PUSHER_STEPS, pusher_model
    g = O.read_vxbin(os.path.join(GOLD, "cilia_pusher_vx3_vx2col.vxbin"))
    m = pusher_model()
    for steps in PUSHER_STEPS:
        w = O.OracleWorld(m)
        w.run(steps)
        s = w.state()
        assert np.abs(s["pos"] - g[f"pos_{steps}"]).max() < 5e-7, steps  # metres, after up to 1.4 m of travel
        assert np.abs(s["orient"] - g[f"orient_{steps}"]).max() < 1e-4, steps
    assert w.result().collision_count > 1000  # the target was being pushed (hit counter, VX3_VoxelyzeKernel.cu:706-708)
    assert np.abs(g["pos_2000"][:, 0]).max() > 1.0


def test_oracle_stop_condition_and_fitness():
    m = gv.Model.from_vxa(W.xenobot_base_vxa(stop_seconds=0.05), W.xenobot_vxd(2))
    w = O.OracleWorld(m)
    n = w.run(1000000, use_stop_formula=True)
    r = w.result()
    assert r.stopped == 1 and r.current_time > 0.05 and r.current_time - r.dt <= 0.05
    assert n == r.steps == int(np.ceil(0.05 / r.dt)) or n == r.steps
    assert np.isfinite(r.fitness_score)
