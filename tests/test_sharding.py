"""Multi-GPU host logic on CPU: worlds are independent, so N ranks only need to agree on who takes which world and to merge
their reports.  world_size-2 gloo processes shard one generation with vx3b_shard_worlds (longest-processing-time-first) and
gather their per-world records; the merged, fitness-sorted list must equal the single-process one."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

import gpuvoxels_b200 as gv
from gpuvoxels_b200 import worlds as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(costs, n):
    costs = np.ascontiguousarray(costs, dtype=np.int64)
    out = np.empty(len(costs), dtype=np.int32)
    rc = gv.lib().vx3b_shard_worlds(costs.ctypes.data_as(C.POINTER(C.c_int64)), len(costs), n, out.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return out


def test_shards_are_a_balanced_partition():
    rng = np.random.default_rng(5)
    costs = rng.integers(80, 2400, size=1024)
    for n in (1, 2, 4, 8):
        s = shard(costs, n)
        assert s.min() == 0 and s.max() == n - 1
        loads = np.array([costs[s == k].sum() for k in range(n)])
        assert loads.sum() == costs.sum()
        assert loads.max() - loads.min() <= costs.max()  # LPT bound
        assert loads.max() <= 1.01 * costs.sum() / n
    assert np.array_equal(shard(costs, 8), shard(costs, 8))  # deterministic: every rank computes the same assignment
    assert len(shard(np.array([], dtype=np.int64), 4)) == 0
    assert gv.lib().vx3b_shard_worlds(None, 3, 0, None) < 0  # bad arguments are errors, not crashes


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = W.xenobot_base_vxa()
    models = [gv.Model.from_vxa(base, W.xenobot_vxd(100 + i)) for i in range(12)]
    costs = np.array([m.num_voxels + 2 * m.num_links for m in models])
    mine = [i for i, s in enumerate(shard(costs, world)) if s == rank]
    # stand-in for ctx.upload(mine); ctx.run(); ctx.results(): a per-world record that depends only on the world itself
    records = [(i, float(models[i].num_voxels) / (1 + models[i].num_links), models[i].num_voxels) for i in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, records)
    merged = sorted((r for part in gathered for r in part), key=lambda r: -r[1])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, merged))


def test_bench_deals_a_generation_to_its_ranks_by_cost(monkeypatch):
    """bench.py --gpus N: every rank computes the same longest-processing-time-first deal (vx3b_shard_worlds on the generator's cost
    estimate) and takes its own share - a disjoint cover of the generation with near-equal cost per rank; N = 1 keeps the plain range"""
    import argparse
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    import bench
    a = argparse.Namespace(workload="xenobot", worlds=24, first_world=7, cube=100, pile=12)
    monkeypatch.setenv("WORLD_SIZE", "1")
    one = bench.world_texts(a, 0)
    assert [d for _, d in one] == [W.xenobot_vxd(7 + i) for i in range(24)]
    monkeypatch.setenv("WORLD_SIZE", "4")
    everything = {W.xenobot_vxd(7 + i): W.xenobot_cost(7 + i) for i in range(96)}
    seen, costs = set(), []
    for rank in range(4):
        mine = [d for _, d in bench.world_texts(a, rank)]
        assert not seen & set(mine)
        seen |= set(mine)
        costs.append(sum(everything[d] for d in mine))
    assert seen == set(everything)
    assert max(costs) - min(costs) <= max(everything.values())  # LPT: no rank is ahead of another by more than one world
    est = [W.xenobot_cost(i) for i in range(40)]
    real = [m.num_voxels + 2 * m.num_links for m in (gv.Model.from_vxa(W.xenobot_base_vxa(), W.xenobot_vxd(i)) for i in range(40))]
    assert np.corrcoef(est, real)[0, 1] > 0.95  # the estimate ranks worlds like their real size does


def test_two_ranks_shard_and_merge_a_generation():
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
    (_, mine0, merged0), (_, mine1, merged1) = outs
    assert sorted(mine0 + mine1) == list(range(12)) and not set(mine0) & set(mine1)
    assert merged0 == merged1 and len(merged0) == 12
    base = W.xenobot_base_vxa()
    single = sorted(((i, float(m.num_voxels) / (1 + m.num_links), m.num_voxels) for i, m in
                     enumerate(gv.Model.from_vxa(base, W.xenobot_vxd(100 + i)) for i in range(12))), key=lambda r: -r[1])
    assert merged0 == single
