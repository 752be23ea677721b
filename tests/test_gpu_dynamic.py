"""SURVEY.md 8f-3 on the GPU: signals, attachment, detachment, removal (k_dynamic, one CTA per world) through the C ABI against
frozen outputs of the reference's OWN Voxelyze3 step code built for the host (tests/golden/vx3ref_*.vxbin, serial thread schedule -
ONE valid execution of kernels that race on a GPU; tests/golden/make_golden_vx3.py) and against the oracle.  Integer work - the grown
link list, removed / detached flags, hit counter - must be equal; the signal scalars are the same double operations on the same
inputs and must be equal wherever the (continuous) state that triggers them has not drifted past a threshold; positions within
the north star's tolerances (BASELINE.json)."""
import os

import numpy as np
import pytest

import gpuvoxels_b200 as gv
import oracle_lib as O
from gpuvoxels_b200 import worlds as W
from golden.make_golden_vx3 import CASES

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
HITS = 4


def gold(name):
    return O.read_vxbin(os.path.join(GOLD, name + ".vxbin"))


def disp_scale(model, pos):
    return max(np.abs(pos - model.array("lattice") * model.desc.vox_size).max(), 1e-30)


def check_against(ctx, world, model, g, sfx, tol, exact_signals=True):
    s = ctx.read_state(world)
    scale = disp_scale(model, g["pos" + sfx])
    err = np.abs(s["pos"] - g["pos" + sfx]).max() / scale
    assert err < tol, ("pos", sfx, err)
    assert np.abs(s["orient"] - g["orient" + sfx]).max() < max(tol * 10, 1e-9), ("orient", sfx)
    if "links" + sfx in g:  # (not kept for the 2,964-voxel world: size)
        topo = ctx.read_topology(world)
        assert np.array_equal(topo[:, :5], g["links" + sfx][:, :5]), ("links", sfx)
        assert np.array_equal(topo[:, 5], g["links" + sfx][:, 5]), ("isNewLink", sfx)
    else:
        assert ctx.num_links(world) == model.num_links  # nothing has attached in the oracle either (tests/test_dynamic_topology.py)
    sig, fl = ctx.read_signals(world)
    assert np.array_equal(fl[:, 0], g["vflags" + sfx][:, 1]), ("removed", sfx)
    if model.desc.enable_attach:
        assert np.array_equal(fl[:, 1], g["vflags" + sfx][:, 2]), ("enableAttach", sfx)
    if exact_signals:
        assert np.array_equal(sig, g["signals" + sfx]), ("signals", sfx, np.abs(sig - g["signals" + sfx]).max())
    return err


DYN = [("vx3ref_signaling_data1_3000", 3000), ("vx3ref_signaling_data2_3000", 3000), ("vx3ref_sticky_pebbles_3000", 3000),
       ("vx3ref_breaking_chain_3000", 3000), ("vx3ref_target_only_300", 300)]


@pytest.mark.parametrize("name,steps", DYN)
def test_fp64_equals_the_reference_step_code(name, steps):
    """fp64 validation build: positions 1e-5 relative (north star), topology / flags / hit counter / signal scalars equal"""
    vxa, vxd, _, checkpoints, _ = CASES[name]
    g = gold(name)
    m = gv.Model.from_vxa(vxa, vxd)
    ctx = gv.Context(0, gv.FP64)
    ctx.upload([m])
    done = 0
    for c, sfx in [(c, f"_{c}") for c in checkpoints] + [(steps, "")]:
        ctx.run(c - done)
        done = c
        err = check_against(ctx, 0, m, g, sfx, 1e-5)
        r = ctx.results()[0]
        assert r.collision_count == g["scalars" + sfx][0, HITS], sfx
        assert r.steps == c
        print(f"{name} step {c}: displacement rel err {err:.2e}, links {ctx.num_links(0)}, hits {r.collision_count}")
    ctx.close()


@pytest.mark.parametrize("name,steps", [("vx3ref_signaling_data1_3000", 1000), ("vx3ref_signaling_data2_3000", 1000), ("vx3ref_sticky_pebbles_3000", 600),
                                        ("vx3ref_breaking_chain_3000", 300)])
def test_fp32_follows_the_reference_step_code(name, steps):
    """production arithmetic over the first checkpoints: the same topology events, positions within 1e-3 of the displacement"""
    vxa, vxd, _, checkpoints, _ = CASES[name]
    g = gold(name)
    m = gv.Model.from_vxa(vxa, vxd)
    ctx = gv.Context(0, gv.FP32)
    ctx.upload([m])
    ctx.run(steps)
    sfx = f"_{steps}"
    s = ctx.read_state(0)
    err = np.abs(s["pos"] - g["pos" + sfx]).max() / disp_scale(m, g["pos" + sfx])
    print(f"{name} fp32 step {steps}: displacement rel err {err:.2e}")
    assert err < 1e-3, err
    topo = ctx.read_topology(0)
    assert np.array_equal(topo[:, :5], g["links" + sfx][:, :5])
    _, fl = ctx.read_signals(0)
    assert np.array_equal(fl[:, 0], g["vflags" + sfx][:, 1])
    ctx.close()


def test_dynamic_and_static_worlds_share_a_context():
    """A batch that mixes static worlds (k_resident) with dynamic ones (k_dynamic): world indices, results and read-backs map to
    the right engine, and every world ends exactly where it ends when run alone."""
    dyn1 = gv.Model.from_vxa(*CASES["vx3ref_sticky_pebbles_3000"][:2])
    dyn2 = gv.Model.from_vxa(*CASES["vx3ref_breaking_chain_3000"][:2])
    st1 = gv.Model.from_vxa(W.xenobot_base_vxa(), W.xenobot_vxd(3))
    st2 = gv.Model.from_vxa(W.small_block_vxa())
    batch = [st1, dyn1, st2, dyn2]
    steps = 400
    alone = []
    for m in batch:
        c = gv.Context(0, gv.FP64)
        c.upload([m])
        c.run(steps)
        alone.append((c.read_state(0), c.results()[0], c.read_topology(0), c.read_signals(0)))
        c.close()
    ctx = gv.Context(0, gv.FP64)
    ctx.upload(batch)
    ctx.run(steps)
    res = ctx.results()
    assert ctx.total_voxel_steps() == steps * sum(m.num_voxels for m in batch)
    for w, m in enumerate(batch):
        s, r, topo, (sig, fl) = alone[w]
        got = ctx.read_state(w)
        for k in ("pos", "orient", "linMom", "angMom"):
            assert np.array_equal(got[k], s[k]), (w, k)
        assert res[w].steps == steps and res[w].num_voxel == m.num_voxels
        # (the static engine's end-of-run centre of mass is an atomic sum: last-bit differences between batch layouts)
        assert np.allclose(res[w].current_com, r.current_com, rtol=1e-12, atol=0) and abs(res[w].fitness_score - r.fitness_score) <= 1e-12 * abs(r.fitness_score)
        if m in (dyn1, dyn2):
            assert list(res[w].current_com) == list(r.current_com) and res[w].fitness_score == r.fitness_score
        assert np.array_equal(ctx.read_topology(w), topo)
        assert np.array_equal(ctx.read_signals(w)[0], sig) and np.array_equal(ctx.read_signals(w)[1], fl)
        assert len(ctx.read_links(w)["flags"]) == ctx.num_links(w)
    assert ctx.num_links(1) > dyn1.num_links  # the pebbles have attached
    # a static world reports its uploaded topology and the initial signal state
    assert np.array_equal(ctx.read_topology(0)[:, 0], st1.array("link_neg")) and np.array_equal(ctx.read_topology(0)[:, 2], st1.array("link_axis"))
    assert not ctx.read_signals(0)[0].any()
    ctx.close()


def test_run_in_pieces_equals_one_run():
    """run(a) + run(b) leaves a dynamic world bit-identical to run(a + b): signal timers, the SafetyGuard countdown of new links
    and the surface list carry over between launches"""
    m = gv.Model.from_vxa(*CASES["vx3ref_sticky_pebbles_3000"][:2])
    whole = gv.Context(0, gv.FP64)
    whole.upload([m])
    whole.run(900)
    parts = gv.Context(0, gv.FP64)
    parts.upload([m])
    for n in (1, 199, 333, 367):
        parts.run(n)
    a, b = whole.read_state(0), parts.read_state(0)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(whole.read_topology(0), parts.read_topology(0))
    assert np.array_equal(whole.read_signals(0)[0], parts.read_signals(0)[0])
    assert whole.results()[0].fitness_score == parts.results()[0].fitness_score
    whole.close(), parts.close()


def test_dynamic_world_is_deterministic_and_matches_the_oracle_pair_set():
    """two runs give the same bits; the watched pair set at a fixed step equals the oracle's (bit-exact collision pair sets)"""
    vxa, vxd = CASES["vx3ref_signaling_data1_3000"][:2]
    m = gv.Model.from_vxa(vxa, vxd)
    w = O.OracleWorld(m)
    w.run(1800)
    states = []
    for _ in range(2):
        ctx = gv.Context(0, gv.FP64)
        ctx.set_option("record_pairs", 1)
        ctx.upload([m])
        ctx.run(1800)
        states.append(ctx.read_state(0))
        pairs = ctx.collision_pairs(0)
        ctx.close()
    for k in states[0]:
        assert np.array_equal(states[0][k], states[1][k]), k
    want = w.collision_pairs()
    assert len(want) > 0 and np.array_equal(np.asarray(pairs).reshape(-1, 2), np.asarray(want).reshape(-1, 2))


def test_material_timers_route_a_world_to_the_dynamic_kernel():
    """TurnOnThermalExpansionAfterThisManySeconds / TurnOnCiliaAfterThisManySeconds (VX3_VoxelyzeKernel.cu:632-634, 852-854) without
    any other dynamic feature: the static kernels do not look at them, so such a world must be stepped by k_dynamic - against the oracle"""
    vxa = W.xenobot_base_vxa().replace("</Mechanical>", "<TurnOnThermalExpansionAfterThisManySeconds>0.02</TurnOnThermalExpansionAfterThisManySeconds>"
                                       "<TurnOnCiliaAfterThisManySeconds>0.01</TurnOnCiliaAfterThisManySeconds></Mechanical>")
    m = gv.Model.from_vxa(vxa, W.xenobot_vxd(5))
    assert any(v.TurnOnCiliaAfterThisManySeconds > 0 for v in m.voxel_materials())
    w = O.OracleWorld(m)
    w.run(600)
    ctx = gv.Context(0, gv.FP64)
    ctx.upload([m])
    ctx.run(600)
    got, ref = ctx.read_state(0)["pos"], w.state()["pos"]
    err = np.abs(got - ref).max() / disp_scale(m, ref)
    print(f"material timers: displacement rel err {err:.2e}, t = {ctx.results()[0].current_time:.4f}")
    assert err < 1e-5
    # and the timers do something: the same world without them ends elsewhere
    m0 = gv.Model.from_vxa(W.xenobot_base_vxa(), W.xenobot_vxd(5))
    c0 = gv.Context(0, gv.FP64)
    c0.upload([m0])
    c0.run(600)
    assert np.abs(c0.read_state(0)["pos"] - got).max() > 1e-6
    ctx.close(), c0.close()


@pytest.mark.parametrize("case", ["c.signaling_data1", "c.signaling_data2"])
def test_job_run_on_the_reference_signaling_experiments(case):
    """Experiments/xenobot/c.signaling/data1|data2 through vx3b_job_run (the drop-in for `Voxelyze3 -i dir -o report.xml`), which
    refused them in round 1: the report's fitness, stop time and centre of mass equal the oracle's run to the stop condition"""
    import tempfile
    import xml.etree.ElementTree as ET
    d = os.path.join(ROOT, "tests", "data", "ref_inputs", case)
    m = gv.Model.from_files(os.path.join(d, "base.vxa"), os.path.join(d, "rob.vxd"))
    w = O.OracleWorld(m)
    w.run(1000000, True)
    want = w.result()
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "report.xml")
        rc = gv.lib().vx3b_job_run(d.encode(), out.encode(), 1, gv.FP64)
        assert rc == 0, gv.lib().vx3b_last_error()
        x = ET.parse(out).getroot().find("detail/rob")
    assert abs(float(x.find("currentTime").text) - want.current_time) < 1e-12
    com = np.array([float(x.find(f"currentCenterOfMass/{a}").text) for a in "xyz"])
    travel = max(np.abs(np.array(want.current_com) - np.array(want.initial_com)).max(), m.desc.vox_size)
    assert np.abs(com - np.array(want.current_com)).max() / travel < 1e-5
    assert abs(float(x.find("fitness_score").text) - want.fitness_score) <= 1e-5 * max(abs(want.fitness_score), 1.0)
