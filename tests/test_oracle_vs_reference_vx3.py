"""Pins the CPU oracle (oracle/vx3_oracle.cpp) against the reference's OWN Voxelyze3 step code built for the host
(oracle/_ref/vx3_ref: Voxelyze3/src/VX3/*.cu compiled where they lie with g++ against oracle/vx3_shim, serial thread schedule):
frozen outputs in tests/golden/vx3ref_*.vxbin (always) and live runs (when the binary has been built from /root/reference).

This is what pins the parts of the VX3 step that Voxelyze2's CPU library cannot: contact pair selection INSIDE a body
(all surface pairs, depth-1 neighbour exclusion: VX3_VoxelyzeKernel.cu:651-727), target hits, the force field
(VX3_Voxel.cu:179-184), the formula machine (Utils/VX3_MathTree.h:50-193), centre-of-mass sampling / heading angle
(VX3_VoxelyzeKernel.cu:314-334) and the fitness (:540-544).  Both sides are compiled without FMA contraction and call the same
libm, so live comparisons are bit for bit; the frozen ones allow 1e-10 relative (another machine's libm may round a sine
differently)."""
import os

import numpy as np
import pytest

import gpuvoxels_b200 as gv
import oracle_lib as O
from gpuvoxels_b200 import worlds as W
from golden.make_golden_vx3 import CASES, CIRCLE_PUSH_BOTS, merged

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SC = {"currentTime": 0, "dt": 1, "OptimalDt": 2, "steps": 3, "collisionCount": 4, "fitness_score": 5, "com": slice(6, 9), "initial_com": slice(9, 12),
      "recentAngle": 12, "targetCloseness": 13, "numClosePairs": 14}


def sorted_pairs(p):
    return np.array(sorted(map(tuple, np.asarray(p).reshape(-1, 2))), dtype=np.int32).reshape(-1, 2)


def check_world(w, ref, sfx, exact, tol=1e-10):
    s, r = w.state(), w.result()
    sc = ref["scalars" + sfx][0]
    for k in ("pos", "orient", "linMom", "angMom"):
        if k + sfx not in ref:
            continue  # (the big fixtures keep positions and orientations only)
        if exact:
            assert np.array_equal(s[k], ref[k + sfx]), (k, sfx, np.abs(s[k] - ref[k + sfx]).max())
        else:
            assert O.rel_err(s[k], ref[k + sfx]) < tol, (k, sfx)
    assert r.steps == sc[SC["steps"]] and r.collision_count == sc[SC["collisionCount"]]
    assert r.dt == sc[SC["dt"]]  # the float-truncated time step, bit for bit
    assert abs(r.current_time - sc[SC["currentTime"]]) <= 1e-15 * max(1.0, abs(r.current_time))
    if "pairs" + sfx in ref:
        assert np.array_equal(sorted_pairs(w.collision_pairs()), ref["pairs" + sfx])  # watched pairs of the last step: exact set
    if "signals" + sfx in ref:  # signal state per voxel, removed / enableAttach flags, the link list as attachment has grown it
        sig, fl = w.signals()
        assert np.abs(sig - ref["signals" + sfx]).max() <= tol * max(1.0, np.abs(ref["signals" + sfx]).max())
        assert np.array_equal(fl[:, 0], ref["vflags" + sfx][:, 1])
        if w.model.desc.enable_attach:  # (without attachment the flag is unobservable - and, with no conditions given, undefined in the reference)
            assert np.array_equal(fl[:, 1], ref["vflags" + sfx][:, 2])
        if "links" + sfx in ref:
            assert np.array_equal(w.topology(), ref["links" + sfx])


FROZEN = [n for n in CASES if not n.endswith("_full")]


@pytest.mark.parametrize("name", FROZEN)
def test_oracle_matches_frozen_vx3_reference_runs(name):
    vxa, vxd, steps, checkpoints, use_stop = CASES[name]
    g = O.read_vxbin(os.path.join(GOLD, name + ".vxbin"))
    m = gv.Model.from_vxa(vxa, vxd)
    w = O.OracleWorld(m)
    done = 0
    # the plastic beams carry amplified y-symmetry noise (tests/test_materials.py) and momenta that pass through zero
    tol = 1e-6 if "plastic" in name else 1e-10
    for c in checkpoints:
        w.run(c - done)
        done = c
        check_world(w, g, f"_{c}", exact=False, tol=tol)
    w.run(steps - done, use_stop_formula=use_stop)
    check_world(w, g, "", exact=False, tol=tol)
    r, sc = w.result(), g["scalars"][0]
    tol = lambda x: 1e-9 * max(1.0, abs(x))  # noqa: E731
    assert abs(r.fitness_score - sc[SC["fitness_score"]]) < tol(sc[SC["fitness_score"]])
    assert np.abs(np.array(r.current_com) - sc[SC["com"]]).max() < 1e-12 and np.abs(np.array(r.initial_com) - sc[SC["initial_com"]]).max() < 1e-15
    if "plastic" not in name:  # (a beam clamped at one end: its centre of mass swings through the same points, the heading between them is noise)
        assert abs(r.recent_angle - sc[SC["recentAngle"]]) < 1e-9


def test_frozen_runs_do_exercise_intra_body_contact_and_target_hits():
    """the fixtures are only worth something if the things they pin actually happen in them"""
    pairs = {n: len(O.read_vxbin(os.path.join(GOLD, n + ".vxbin")).get("pairs", ())) for n in FROZEN}
    hits = {n: O.read_vxbin(os.path.join(GOLD, n + ".vxbin"))["scalars"][0, SC["collisionCount"]] for n in FROZEN}
    assert pairs["vx3ref_xeno7_contact_500"] > 20 and pairs["vx3ref_xeno15_contact_stop1s"] > 50 and pairs["vx3ref_pile_k3c4_600"] > 100
    assert hits["vx3ref_xeno15_contact_stop1s"] > 1000 and hits["vx3ref_circle_push_bot_0003_1000"] > 100
    assert sum(1 for n in FROZEN if n.startswith("vx3ref_circle_push") and pairs[n] > 0) >= 3  # real bots touch themselves too
    # xeno7: every watched pair is INSIDE the one body (it has no pebble contact yet) - voxels of the same connected component
    m = gv.Model.from_vxa(W.xenobot_base_vxa(), W.xenobot_vxd(7))
    mats = m.array("voxel_mat")
    p = O.read_vxbin(os.path.join(GOLD, "vx3ref_xeno7_contact_500.vxbin"))["pairs"]
    assert np.all(mats[p[:, 0]] == mats[p[:, 1]])


@pytest.mark.skipif(not O.have_reference_vx3(), reason="oracle/_ref/vx3_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name,steps", [("vx3ref_xeno7_contact_500", 400), ("vx3ref_forcefield_800", 300), ("vx3ref_pile_k3c4_600", 300),
                                        ("vx3ref_circle_push_bot_0009_1000", 700), ("vx3ref_circle_push_bot_0007_1000", 1500)])
def test_oracle_is_bit_identical_to_the_live_vx3_reference(name, steps):
    vxa, vxd, _, _, _ = CASES[name]
    ref, _ = O.run_reference_vx3(merged(vxa, vxd), steps, dump_at=(1, steps // 2))
    w = O.OracleWorld(gv.Model.from_vxa(vxa, vxd))
    done = 0
    for c, sfx in ((1, "_1"), (steps // 2, f"_{steps // 2}"), (steps, "")):
        w.run(c - done)
        done = c
        check_world(w, ref, sfx, exact=True)


@pytest.mark.skipif(not O.have_reference_vx3(), reason="oracle/_ref/vx3_ref not built (needs /root/reference)")
def test_whole_run_to_the_stop_condition_is_bit_identical():
    """a world from the reference's own experiment, run to its stop formula (t = 5 s, 9,935 steps): step count, hit counter,
    heading angle, centre of mass and the fitness the report would carry are the reference's to the last bit"""
    vxa, vxd, steps, _, _ = CASES["vx3ref_circle_push_bot_0006_full"]
    ref, _ = O.run_reference_vx3(merged(vxa, vxd), steps, use_stop=True)
    w = O.OracleWorld(gv.Model.from_vxa(vxa, vxd))
    w.run(steps, use_stop_formula=True)
    r, sc = w.result(), ref["scalars"][0]
    assert r.steps == sc[SC["steps"]] == 9935
    assert r.fitness_score == sc[SC["fitness_score"]] and r.recent_angle == sc[SC["recentAngle"]]
    assert np.array_equal(np.array(r.current_com), sc[SC["com"]])
    assert np.array_equal(w.state()["pos"], ref["pos"])


@pytest.mark.parametrize("bot", CIRCLE_PUSH_BOTS)
def test_host_model_of_the_reference_experiment_inputs(bot):
    """the reference's e.circle_push VXDs load, and the model is the one its own reader builds (voxel count of the frozen run)"""
    vxa, vxd, _, _, _ = CASES[f"vx3ref_circle_push_{bot}_1000"]
    m = gv.Model.from_vxa(vxa, vxd)
    g = O.read_vxbin(os.path.join(GOLD, f"vx3ref_circle_push_{bot}_1000.vxbin"))
    assert m.num_voxels == len(g["pos"])
    assert m.desc.enable_cilia == 1 and m.desc.enable_collision == 1
