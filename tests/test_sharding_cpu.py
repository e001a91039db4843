"""CPU: the multi-rank host logic over gloo (world_size 2) and in lockstep emulation - slide assignment,
equal-count strips, the all-to-all partition, and the halo exchange protocol (counts first, NaN-padded payload).
The CUDA pack / partition kernels are replaced here by torch-CPU stand-ins with the same contract (test doubles only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from path_gene_multimodal_b200 import sharding


class FakeEngine:
    """CPU stand-in for Engine.halo_pack / check_overflow (same record layout: x, y, {gid, type})."""

    def halo_pack(self, xy, types, gid, lo_edge, hi_edge, capacity):
        take = (xy[:, 0] < lo_edge) | (xy[:, 0] >= hi_edge)
        recs = torch.zeros((max(capacity, 1), 3), dtype=torch.float64)
        n = int(take.sum())
        recs[:n, :2] = xy[take]
        recs[:n, 2] = torch.stack([gid[take].to(torch.int32), types[take].to(torch.int32)], dim=1).contiguous() \
            .view(torch.float64).reshape(-1)
        return recs, torch.tensor([n], dtype=torch.int32)

    def check_overflow(self):
        pass

    def strip_partition(self, xy, types, gid, inner_edges):
        """Stand-in for pg_strip_partition: records grouped by owning strip (input order kept), totals per strip."""
        inner = torch.as_tensor(np.asarray(inner_edges, dtype=np.float64))
        owner = torch.bucketize(xy[:, 0].contiguous(), inner, right=True)
        order = torch.argsort(owner, stable=True)
        rec = torch.empty((xy.shape[0], 3), dtype=torch.float64)
        rec[:, :2] = xy
        rec[:, 2] = torch.stack([gid.to(torch.int32), types.to(torch.int32)], dim=1).contiguous().view(torch.float64).reshape(-1)
        return rec[order].contiguous(), torch.bincount(owner, minlength=len(inner_edges) + 1).to(torch.int32)


def _points(n, seed):
    rng = np.random.default_rng(seed)
    xy = torch.from_numpy(rng.random((n, 2)) * 1000.0)
    ty = torch.from_numpy(rng.integers(1, 6, size=n).astype(np.int32))
    return xy, ty


def test_assign_slides():
    assert sharding.assign_slides(5, 2) == [[0, 2, 4], [1, 3]]
    a = sharding.assign_slides(6, 3, sizes=[10, 1, 1, 1, 5, 5])
    assert sorted(sum(a, [])) == list(range(6))
    loads = [sum([10, 1, 1, 1, 5, 5][s] for s in r) for r in a]
    assert max(loads) == 10 and min(loads) >= 6
    assert sharding.slide_parallel(3, lambda s: s * s) == {0: 0, 1: 1, 2: 4}


def test_strips_and_partition_emulated():
    world = 4
    xy, ty = _points(4000, 1)
    gid = torch.arange(4000, dtype=torch.int32)
    chunks = [slice(q * 1000, (q + 1) * 1000) for q in range(world)]
    edges = sharding.run_emulated([sharding.equal_count_edges(xy[c][:, 0], world, 0.0, 1000.0) for c in chunks])
    assert all(np.array_equal(edges[0], e) for e in edges)
    e = edges[0]
    assert e[0] == 0.0 and e[-1] == 1000.0 and np.all(np.diff(e) > 0)
    cnt = np.histogram(xy[:, 0].numpy(), bins=e)[0]
    assert cnt.sum() == 4000 and abs(cnt - 1000).max() < 40        # equal counts up to one histogram bin
    parts = sharding.run_emulated([sharding.partition_by_strips(FakeEngine(), xy[c], ty[c], gid[c], e, q, world) for q, c in enumerate(chunks)])
    seen = []
    for q, (pxy, pty, pgid) in enumerate(parts):
        assert bool(((pxy[:, 0] >= e[q]) & ((pxy[:, 0] < e[q + 1]) | (q == world - 1))).all())
        assert torch.equal(pxy, xy[pgid.long()]) and torch.equal(pty, ty[pgid.long()])   # records travel intact
        seen.append(pgid)
    assert torch.equal(torch.sort(torch.cat(seen)).values, gid)


def test_halo_exchange_emulated():
    world = 3
    xy, ty = _points(3000, 2)
    gid = torch.arange(3000, dtype=torch.int32)
    strips = sharding.strips_from_edges([0.0, 300.0, 650.0, 1000.0])
    own = [((xy[:, 0] >= s.x_lo) & (xy[:, 0] < s.x_hi)) for s in strips]
    eng = FakeEngine()
    res = sharding.run_emulated([sharding.exchange_halo(eng, xy[m], ty[m], gid[m], s, 40.0, q, world)
                                 for q, (s, m) in enumerate(zip(strips, own))])
    for q, (all_recs, max_cnt) in enumerate(res):
        assert all_recs.shape == (world * max_cnt, 3)
        valid = ~torch.isnan(all_recs[:, 0])
        x = all_recs[valid, 0]
        meta = all_recs[valid, 2].contiguous().view(torch.int32).reshape(-1, 2)
        # every point within 40 px of an interior strip edge was shipped exactly once, with its gid and type
        near = torch.zeros(3000, dtype=torch.bool)
        for s in strips:
            if not s.is_first:
                near |= (xy[:, 0] >= s.lo) & (xy[:, 0] < s.lo + 40.0)
            if not s.is_last:
                near |= (xy[:, 0] >= s.hi - 40.0) & (xy[:, 0] < s.hi)
        assert torch.equal(torch.sort(meta[:, 0]).values, gid[near])
        assert torch.equal(x, xy[meta[:, 0].long(), 0]) and torch.equal(meta[:, 1], ty[meta[:, 0].long()])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = sharding.TorchComm()
        xy, ty = _points(4000, 1)
        gid = torch.arange(4000, dtype=torch.int32)
        mine = slice(rank * 2000, (rank + 1) * 2000)          # each rank starts with an arbitrary half of the table
        edges = sharding.run(sharding.equal_count_edges(xy[mine][:, 0], world, 0.0, 1000.0), comm)
        pxy, pty, pgid = sharding.run(sharding.partition_by_strips(FakeEngine(), xy[mine], ty[mine], gid[mine], edges, rank, world), comm)
        strip = sharding.strips_from_edges(edges)[rank]
        ok = bool(((pxy[:, 0] >= strip.x_lo) & (pxy[:, 0] < strip.x_hi)).all()) and torch.equal(pxy, xy[pgid.long()])
        all_recs, max_cnt = sharding.run(sharding.exchange_halo(FakeEngine(), pxy, pty, pgid, strip, 25.0, rank, world), comm)
        other = all_recs[(1 - rank) * max_cnt:(2 - rank) * max_cnt]
        other = other[~torch.isnan(other[:, 0])]
        lo, hi = (edges[1] - 25.0, edges[1]) if rank == 1 else (edges[1], edges[1] + 25.0)
        expect = int(((xy[:, 0] >= lo) & (xy[:, 0] < hi)).sum())
        ok = ok and other.shape[0] == expect
        stats = sharding.slide_parallel(5, lambda s: {"slide": s, "rank": rank}, rank, world)
        ok = ok and sorted(stats) == [0, 1, 2, 3, 4] and stats[3]["rank"] == 1
        q.put((rank, ok, int(pxy.shape[0]), [float(e) for e in edges]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gloo_world2_partition_and_halo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    out.sort()
    assert all(o[1] for o in out), out
    assert out[0][2] + out[1][2] == 4000 and abs(out[0][2] - 2000) < 40 and out[0][3] == out[1][3]
