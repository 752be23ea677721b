"""One body over several GPUs (gpuvoxels_b200/slab.py): the slab decomposition and its exchange plan on the
CPU - invariants of split_body, and two gloo ranks that really exchange their boundary lists - and, on the GPU, slabs stepped as
separate contexts with halo exchange against the same body stepped whole."""
import os
import socket

import numpy as np
import pytest

import gpuvoxels_b200 as gv
from gpuvoxels_b200 import slab
from gpuvoxels_b200 import worlds as W


def bodies():
    return {"cube7": gv.Model.from_vxa(W.cube_vxa(7)), "xenobot": gv.Model.from_vxa(W.xenobot_base_vxa(enable_collision=False), W.xenobot_vxd(9)),
            "block": gv.Model.from_vxa(W.small_block_vxa())}


@pytest.mark.parametrize("name", ["cube7", "xenobot", "block"])
@pytest.mark.parametrize("n_parts,axis", [(2, 2), (3, 2), (2, 0)])
def test_split_body_invariants(name, n_parts, axis):
    model = bodies()[name]
    lat = model.array("lattice")
    if lat[:, axis].max() - lat[:, axis].min() + 1 < n_parts:
        pytest.skip("fewer layers than slabs")
    parts = slab.split_body(model, n_parts, axis)
    nv, nl = model.num_voxels, model.num_links
    owner = np.full(nv, -1)
    link_seen = np.zeros(nl, dtype=int)
    full_links = {(a, b): i for i, (a, b) in enumerate(zip(model.array("link_neg"), model.array("link_pos")))}
    for p in parts:
        g = p.global_index
        assert np.all(np.diff(g) > 0)  # creation order kept
        assert np.all(owner[g[p.own]] == -1)
        owner[g[p.own]] = p.rank
        assert np.array_equal(p.model.array("ghost").astype(bool), ~p.own)
        assert np.array_equal(p.model.array("lattice"), lat[g]) and np.array_equal(p.model.array("voxel_mat"), model.array("voxel_mat")[g])
        sn, sp = p.model.array("link_neg"), p.model.array("link_pos")
        assert np.all(p.own[sn] | p.own[sp])  # no ghost-ghost link
        for a, b, ax, lm in zip(g[sn], g[sp], p.model.array("link_axis"), p.model.array("link_mat")):
            i = full_links[(a, b)]
            link_seen[i] += 1
            assert model.array("link_axis")[i] == ax and model.array("link_mat")[i] == lm
        # every link of an own voxel is present: its degree in the slab equals its degree in the body
        deg_full = np.bincount(np.concatenate([model.array("link_neg"), model.array("link_pos")]), minlength=nv)
        deg_sub = np.bincount(np.concatenate([sn, sp]), minlength=len(g))
        assert np.array_equal(deg_sub[p.own], deg_full[g[p.own]])
        assert set(p.ghost_lo) | set(p.ghost_hi) == set(np.nonzero(~p.own)[0]) and not set(p.ghost_lo) & set(p.ghost_hi)
    assert np.all(owner >= 0)
    cut = owner[model.array("link_neg")] != owner[model.array("link_pos")]
    assert np.array_equal(link_seen, 1 + cut.astype(int))  # a cut link lives on both sides, every other link once
    assert parts[0].ghost_lo.size == 0 and parts[-1].ghost_hi.size == 0
    dt = slab.body_time_step(model)
    assert dt > 0 and dt == np.float32(dt)


def _body_texts():
    return {"cube7": (W.cube_vxa(7), None), "xenobot": (W.xenobot_base_vxa(enable_collision=False), W.xenobot_vxd(9)), "block": (W.small_block_vxa(), None)}


@pytest.mark.parametrize("name", ["cube7", "xenobot", "block"])
@pytest.mark.parametrize("n_parts,axis", [(2, 2), (3, 2), (2, 0), (2, 1)])
def test_native_slab_model_equals_the_numpy_split(name, n_parts, axis):
    """vx3b_model_from_vxa_slab (C++: every rank parses the text and builds ONLY its own slab) against split_body (numpy, on the
    whole body's model): same voxels in the same order, same ghosts, links, exchange lists, materials and the whole body's time step"""
    vxa, vxd = _body_texts()[name]
    model = gv.Model.from_vxa(vxa, vxd)
    lat = model.array("lattice")
    if lat[:, axis].max() - lat[:, axis].min() + 1 < n_parts:
        with pytest.raises(gv.VX3Error):
            gv.Model.from_vxa_slab(vxa, vxd, axis, 0, n_parts)
        return
    parts = slab.split_body(model, n_parts, axis)
    vm_whole, lm_whole = model.voxel_materials(), model.link_materials()
    for p in parts:
        m = gv.Model.from_vxa_slab(vxa, vxd, axis, p.rank, n_parts)
        si = m.slab_info()
        assert (si["part"], si["n_parts"], si["axis"], si["num_own"]) == (p.rank, n_parts, axis, p.num_own)
        assert si["body_voxels"] == model.num_voxels and si["body_links"] == model.num_links
        assert si["body_dt"] == slab.body_time_step(model)
        assert np.array_equal(si["global_index"], p.global_index)
        for k in ("send_lo", "send_hi", "ghost_lo", "ghost_hi"):
            assert np.array_equal(si[k], getattr(p, k)), k
        assert m.num_voxels == p.model.num_voxels and m.num_links == p.model.num_links
        for k in ("lattice", "ghost", "link_neg", "link_pos", "link_axis", "phase_offset", "base_cilia", "shift_cilia"):
            assert np.array_equal(m.array(k), p.model.array(k)), k
        # the combined (link) materials are the WHOLE body's, numbered and oriented as the whole body first uses them
        assert np.array_equal(m.array("voxel_mat"), p.model.array("voxel_mat")) and np.array_equal(m.array("link_mat"), p.model.array("link_mat"))
        assert [bytes(x) for x in m.voxel_materials()] == [bytes(x) for x in vm_whole]
        assert [bytes(x) for x in m.link_materials()] == [bytes(x) for x in lm_whole]
        for f in ("dt_frac", "vox_size", "floor_enabled", "vary_temp_enabled", "temp_amplitude", "temp_period", "vx2_compat", "enable_cilia"):
            assert getattr(m.desc, f) == getattr(model.desc, f), f


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = gv.Model.from_vxa(W.cube_vxa(6))
    part = slab.split_body(model, world)[rank]
    lat = model.array("lattice")
    # the record a rank would pack for a boundary voxel: here its global number and lattice cell
    rec = lambda idx: torch.from_numpy(np.concatenate([part.global_index[idx][:, None], lat[part.global_index[idx]]], axis=1).astype(np.int64))
    got = {}
    ops = []
    if rank > 0:
        got["lo"] = torch.empty((len(part.ghost_lo), 4), dtype=torch.int64)
        ops += [dist.P2POp(dist.isend, rec(part.send_lo), rank - 1), dist.P2POp(dist.irecv, got["lo"], rank - 1)]
    if rank < world - 1:
        got["hi"] = torch.empty((len(part.ghost_hi), 4), dtype=torch.int64)
        ops += [dist.P2POp(dist.isend, rec(part.send_hi), rank + 1), dist.P2POp(dist.irecv, got["hi"], rank + 1)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    ok = True
    for side, idx in (("lo", part.ghost_lo), ("hi", part.ghost_hi)):
        if side in got:
            ok = ok and torch.equal(got[side], rec(idx)) and len(idx) == 36
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok), part.num_own, int((~part.own).sum())))


def test_two_gloo_ranks_exchange_their_boundary_layers():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs == [(0, True, 108, 36), (1, True, 108, 36)]


def run_slabs(model, n_parts, steps, precision, options=None):
    import torch
    parts = slab.split_body(model, n_parts)
    dt = slab.body_time_step(model)
    stream = torch.cuda.Stream(0)
    runners = [slab.SlabRunner(p, 0, precision, dt, options, stream=stream) for p in parts]
    with torch.cuda.stream(stream):
        for _ in range(steps):
            for r in runners:
                r.step_begin()
            for lo, hi in zip(runners[:-1], runners[1:]):  # same device, same stream: the exchange is two device copies per cut
                hi.buf[slab.GHOST_LO].copy_(lo.buf[slab.SEND_HI])
                lo.buf[slab.GHOST_HI].copy_(hi.buf[slab.SEND_LO])
            for r in runners:
                r.step_end()
    torch.cuda.synchronize()
    nv = model.num_voxels
    out = {"pos": np.zeros((nv, 3)), "orient": np.zeros((nv, 4)), "linMom": np.zeros((nv, 3)), "angMom": np.zeros((nv, 3))}
    for r in runners:
        st, g = r.own_state()
        for k in out:
            out[k][g] = st[k]
        assert r.ctx.results()[0].steps == steps
        r.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [gv.FP64, gv.FP32])
@pytest.mark.parametrize("n_parts", [2, 3])
def test_slabs_with_halo_exchange_equal_the_whole_body_bit_for_bit(precision, n_parts):
    """Every voxel and every link is stepped by the same code from the same inputs whichever slab it lands in, so the sliced
    run must reproduce the whole-body run exactly - in both precisions - including the cut links' history on both sides."""
    steps = 150
    for model, opts in ((gv.Model.from_vxa(W.cube_vxa(9)), dict(brick_x=4, brick_y=3, brick_z=2)),
                        (gv.Model.from_vxa(W.xenobot_base_vxa(enable_collision=False), W.xenobot_vxd(9)), None)):
        whole = gv.Context(0, precision)
        whole.set_option("path", 5)
        for k, v in (opts or {}).items():
            whole.set_option(k, v)
        whole.upload([model])
        whole.run(steps)
        ref = whole.read_state(0)
        whole.close()
        got = run_slabs(model, n_parts, steps, precision, opts)
        for k in ("pos", "orient", "linMom", "angMom"):
            assert np.array_equal(got[k], ref[k]), (k, np.abs(got[k] - ref[k]).max())


def run_slabs_device_exchange(vxa, vxd, n_parts, steps, precision, options=None, calls=(None,), fused=True):
    """slabs built by the C++ host model, one context + one stream each on device 0, exchanging through mapped receive areas"""
    import torch
    parts = [slab.native_part(vxa, vxd, r, n_parts) for r in range(n_parts)]
    runners = [slab.SlabRunner(p, 0, precision, p.body_dt, options, stream=torch.cuda.Stream(0)) for p in parts]
    eps = [r.open_endpoints() for r in runners]
    for k, r in enumerate(runners):
        r.connect(eps[k - 1]["hi"] if k > 0 else None, eps[k + 1]["lo"] if k < n_parts - 1 else None, same_process=True, fused=fused)
    for n in calls:  # several calls: the sequence numbers carry on from one to the next
        for r in runners:
            r.run(steps if n is None else n)
    torch.cuda.synchronize()
    nv = parts[0].body_voxels
    out = {"pos": np.zeros((nv, 3)), "orient": np.zeros((nv, 4)), "linMom": np.zeros((nv, 3)), "angMom": np.zeros((nv, 3))}
    for r in runners:
        assert r.ctx.halo_timeouts() == 0
        st, g = r.own_state()
        for k in out:
            out[k][g] = st[k]
        assert r.ctx.results()[0].steps == (steps if calls == (None,) else sum(calls))
        r.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("precision", [gv.FP64, gv.FP32])
@pytest.mark.parametrize("n_parts", [2, 3])
def test_device_side_exchange_equals_the_whole_body_bit_for_bit(precision, n_parts, fused):
    """vx3b_batch_run_slab, 64 steps per graph, no host in the loop - against the same body stepped whole; slabs from
    vx3b_model_from_vxa_slab.  fused: k_brick itself stores the boundary voxels into the neighbour's mapped state arrays and raises
    its flag (one launch per time step); staged: k_halo_push into the neighbour's receive area + flag, k_halo_pull on the other side"""
    for (vxa, vxd), opts, calls in (((W.cube_vxa(9), None), dict(brick_x=4, brick_y=3, brick_z=2), (150,)),
                                   ((W.xenobot_base_vxa(enable_collision=False), W.xenobot_vxd(9)), None, (70, 1, 79))):
        model = gv.Model.from_vxa(vxa, vxd)
        whole = gv.Context(0, precision)
        whole.set_option("path", 5)
        for k, v in (opts or {}).items():
            whole.set_option(k, v)
        whole.upload([model])
        whole.run(150)
        ref = whole.read_state(0)
        whole.close()
        got = run_slabs_device_exchange(vxa, vxd, n_parts, 150, precision, opts, calls, fused)
        for k in ("pos", "orient", "linMom", "angMom"):
            assert np.array_equal(got[k], ref[k]), (k, np.abs(got[k] - ref[k]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_a_neighbour_that_never_steps_is_reported_not_hung_on(fused):
    """the wait for a neighbour's flag gives up after 10 s - once: later launches do not wait again - and the run is flagged invalid"""
    import time
    import torch
    vxa = W.cube_vxa(6)
    parts = [slab.native_part(vxa, None, r, 2) for r in range(2)]
    runners = [slab.SlabRunner(p, 0, gv.FP32, p.body_dt, None, stream=torch.cuda.Stream(0)) for p in parts]
    eps = [r.open_endpoints() for r in runners]
    runners[0].connect(None, eps[1]["lo"], same_process=True, fused=fused)
    runners[1].connect(eps[0]["hi"], None, same_process=True, fused=fused)
    t0 = time.time()
    runners[0].run(5)  # slab 1 never runs
    torch.cuda.synchronize()
    assert 8 < time.time() - t0 < 45  # one 10 s wait, not one per step
    assert runners[0].ctx.halo_timeouts() >= 1
    assert runners[0].ctx.results()[0].steps == 5
    for r in runners:
        r.close()


@pytest.mark.gpu
def test_asynchronous_runs_keep_state_in_brick_order_until_read_back():
    """vx3b_batch_run_async leaves the state in the engine's internal order; any read-back must synchronise it, and mixing
    asynchronous and synchronous runs must give what one synchronous run gives."""
    model = gv.Model.from_vxa(W.cube_vxa(6))
    a, b = gv.Context(0, gv.FP64), gv.Context(0, gv.FP64)
    for c in (a, b):
        c.set_option("path", 5)
        c.upload([model])
    a.run(90)
    b.run_async(20)
    b.run_async(30)
    assert b.results()[0].steps == 50
    b.run(25)
    b.run_async(15)
    sa, sb = a.read_state(0), b.read_state(0)
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    ra, rb = a.results()[0], b.results()[0]
    assert ra.steps == rb.steps == 90 and ra.current_time == rb.current_time
    assert np.allclose(ra.current_com, rb.current_com, rtol=1e-12, atol=0)
    with pytest.raises(gv.VX3Error):
        c = gv.Context(0, gv.FP64)
        c.upload([model])  # automatic path: a body this small is stepped by k_resident, which run_async does not drive
        c.run_async(5)
    a.close()
    b.close()
