"""Real multi-GPU check (not collected by pytest): torchrun --nproc-per-node N tests/multi_gpu_check.py
Each rank owns one x-strip of a synthetic slide, exchanges halos over NCCL and rank 0 verifies the concatenated
result against a single-GPU build of the whole slide, bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from path_gene_multimodal_b200 import sharding, synth  # noqa: E402
from path_gene_multimodal_b200.engine import default_knn_cell, get_engine, radius_cell  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = sharding.TorchComm()
    eng = get_engine(local)
    n, k = int(os.environ.get("PG_N", 2_000_000)), 16
    xy, ty, side = synth.make_points(n, 1005)
    # the table starts row-partitioned (e.g. parquet row groups); strips by equal counts, then one all-to-all
    rows = slice(rank * n // world, (rank + 1) * n // world)
    l_xy, l_ty = torch.from_numpy(xy[rows]).to(dev), torch.from_numpy(ty[rows]).to(dev)
    l_gid = torch.arange(rows.start, rows.stop, dtype=torch.int32, device=dev)
    edges = sharding.run(sharding.equal_count_edges(l_xy[:, 0].contiguous(), world, 0.0, float(side)), comm)
    s_xy, s_ty, s_gid = sharding.run(sharding.partition_by_strips(eng, l_xy, l_ty, l_gid, edges, rank, world), comm)
    strip = sharding.strips_from_edges(edges)[rank]
    bounds = (0.0, 0.0, float(side), float(side))
    if os.environ.get("PG_C5_SORT", "1") != "0":   # a strip owns its row order: keep it in cell order (sharding.spatial_sort)
        s_xy, s_ty, s_gid = sharding.spatial_sort(eng, s_xy, s_ty, s_gid, radius_cell(50.0), bounds)
    # PG_PEER=1: the halo exchange as one pack+store kernel over NVLink peer memory instead of the NCCL all-gathers
    peer = sharding.PeerHalo(int(os.environ.get("PG_PEER_CAP", 262144)), dev) if os.environ.get("PG_PEER") == "1" else None
    import time

    def timed(make, reps=int(os.environ.get("PG_REPS", 3))):
        """max-over-ranks wall time (barrier + synchronize on both sides) of the sharded build, best of `reps` after one
        untimed run (workspace growth); bench.py's c5_strip_sharded stage times the same builds with CUDA events."""
        best, out = None, None
        sharding.run(make(), comm)
        for _ in range(reps):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = sharding.run(make(), comm)
            torch.cuda.synchronize(); dist.barrier()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            best = float(dt.item()) if best is None else min(best, float(dt.item()))
        return out, best

    rg, t_rad = timed(lambda: sharding.sharded_radius_graph(eng, s_xy, s_ty, s_gid, 50.0, strip, rank, world, bounds=bounds, peer=peer))
    kg, t_knn = timed(lambda: sharding.sharded_knn_graph(eng, s_xy, s_ty, s_gid, k, strip, rank, world, n_global=n, bounds=bounds, peer=peer))
    if rank == 0:
        print(f"sharded build, world={world} n={n} halo={'nvlink peer memory' if peer else 'nccl'}: radius r=50 {t_rad * 1e3:.2f} ms ({n / t_rad / 1e6:.0f} M nuclei/s), "
              f"kNN k={k} + union + edges {t_knn * 1e3:.2f} ms ({n / t_knn / 1e6:.0f} M nuclei/s)  [halo exchange included]")
    # gather to rank 0 for the comparison
    def gather(t):
        sizes = [None] * world
        dist.all_gather_object(sizes, tuple(t.shape))
        if rank == 0:
            outs = [torch.empty(s, dtype=t.dtype, device=dev) for s in sizes]
            outs[0] = t
            for q in range(1, world):
                dist.recv(outs[q], src=q)
            return outs
        dist.send(t.contiguous(), dst=0)
        return None
    parts = {name: gather(t) for name, t in (("gid", s_gid), ("r_edges", rg["edges"]), ("r_deg", rg["degree"]),
                                             ("k_idx", kg["knn_idx"]), ("k_dist", kg["dist"]), ("k_edges", kg["edges"]),
                                             ("k_deg", kg["degree"]))}
    if rank == 0:
        d_xy, d_ty = torch.from_numpy(xy).to(dev), torch.from_numpy(ty).to(dev)
        eng.grid_build(d_xy, d_ty, None, radius_cell(50.0), None)
        ref = eng.radius_graph(50.0, upper=True, want_edges=True)
        e = torch.cat(parts["r_edges"])
        key = e[:, 0] * n + e[:, 1]
        assert torch.equal(e[torch.argsort(key)], ref["edges"]), "radius edges differ"
        gid = torch.cat(parts["gid"]).long()
        deg = torch.empty(n, dtype=torch.int32, device=dev)
        deg[gid] = torch.cat(parts["r_deg"])
        assert torch.equal(deg, ref["degree"]), "radius degrees differ"
        eng.grid_build(d_xy, d_ty, None, default_knn_cell(n, float(side) ** 2, k), None)
        kn = eng.knn(k, dist_dtype=torch.float64)
        idx = torch.empty((n, k), dtype=torch.int32, device=dev)
        idx[gid] = torch.cat(parts["k_idx"])
        dd = torch.empty((n, k), dtype=torch.float64, device=dev)
        dd[gid] = torch.cat(parts["k_dist"])
        assert torch.equal(idx, kn["knn_idx"]) and torch.equal(dd, kn["dist"]), "kNN lists differ"
        sym = eng.symmetrize(kn["knn_idx"], kn["dist"])
        up = eng.csr_upper(sym["row_ptr"], sym["col"], sym["w64"])
        e = torch.cat(parts["k_edges"])
        key = e[:, 0] * n + e[:, 1]
        assert torch.equal(e[torch.argsort(key)], up["edges"]), "kNN union edges differ"
        print(f"multi_gpu_check ok: world={world} n={n} radius_edges={len(ref['edges'])} knn_union_edges={len(up['edges'])} "
              f"halo={kg['halo']:.1f}px ghosts(rank0)={kg['n_ghost']}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
