"""C4 (BASELINE configs[3]): a TNBC-cohort-sized synthetic batch, slide-parallel (not collected by pytest).

    python tests/cohort_check.py                                   # 1 GPU
    torchrun --nproc-per-node 8 tests/cohort_check.py              # 8 GPUs, slides dealt out by nuclei count

Every slide (PG_SLIDE_N nuclei, seed 1004 + slide index, ragged float32 rings) goes through the whole nuclei-table
pass from HOST arrays (page-locked): H2D of the table, tile->WSI map + morphology, kNN k=8 + undirected union + i<j edges +
composition, radius r=50 px graph with composition / degree statistics, D2H of per-slide summaries only (the
graphs stay on the GPU that built them, as each LSF job of the reference keeps its own slide). PG_LANES slides
(default 2) are in flight per GPU, each on its own stream / handle / host thread, so one slide's H2D overlaps the
other's kernels. Rank 0 prints one
line with the max-over-ranks wall time and a checksum that must not depend on the number of GPUs."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from path_gene_multimodal_b200 import _host, sharding, synth  # noqa: E402
from path_gene_multimodal_b200.engine import Engine, default_knn_cell, get_engine, radius_cell  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    n_slides = int(os.environ.get("PG_SLIDES", 64))
    n = int(os.environ.get("PG_SLIDE_N", 500_000))
    lanes = int(os.environ.get("PG_LANES", 2))   # slides in flight per GPU: one engine (handle) + stream + host thread each
    engines = [get_engine(local)] + [Engine(local) for _ in range(lanes - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    mine = sharding.assign_slides(n_slides, world, sizes=[n] * n_slides)[rank]
    # host tables in page-locked memory, as a loader that reads Parquet straight into pinned buffers would leave
    # them (generation and pinning untimed); _host.to_device then DMAs from them without a staging copy
    def pin(a):
        buf = _host.pinned_empty(a.shape, a.dtype)
        buf[...] = a
        return buf

    tables = {}
    for s in mine:
        tab = synth.make_table(n, synth.SEEDS["C4"] + s)
        for name in ("poly_off", "poly_xy", "nuc_tile", "tile_x", "tile_y", "centroid", "bbox", "types"):
            setattr(tab, name, pin(getattr(tab, name)))
        tables[s] = tab

    def one_slide(s, eng):
        tab = tables[s]
        side_px = float(tab.n_tiles_side * 508)
        t_off = _host.to_device(tab.poly_off, np.int32, dev)
        t_xy = _host.to_device(tab.poly_xy, np.float32, dev)
        t_tile = _host.to_device(tab.nuc_tile, np.int32, dev)
        t_tx, t_ty = _host.to_device(tab.tile_x, np.int32, dev), _host.to_device(tab.tile_y, np.int32, dev)
        t_cen, t_bb = _host.to_device(tab.centroid, np.float64, dev), _host.to_device(tab.bbox, np.int32, dev)
        t_types = _host.to_device(tab.types, np.int32, dev)
        mm = eng.map_morph(t_off, t_xy, t_tile, t_tx, t_ty, t_cen, t_bb, write_polygons=True)
        wsi = mm["wsi_centroid"]
        bnd = (0.0, 0.0, side_px, side_px)
        eng.grid_build(wsi, t_types, None, default_knn_cell(n, side_px ** 2, 8), bnd)
        kn = eng.knn(8, dist_dtype=torch.float32)
        up = comp = eng.knn_union(kn["knn_idx"], kn["dist32"], types=t_types, n_types=5, symmetric_dist=True)
        eng.grid_build(wsi, t_types, None, radius_cell(50.0), bnd)
        rg = eng.radius_graph(50.0, upper=True, n_types=5, want_dist32=True, want_edges=True)
        st = eng.decode_stats(rg["stats"], rg["hist"])
        # small per-slide summary (what a cohort table would keep); the reads synchronise the slide
        return {"knn_edges": int(up["edges"].shape[0]), "radius_edges": int(rg["edges"].shape[0]),
                "area_sum": float(mm["area"].double().sum().item()), "knn_deg_sum": int(comp["degree"].sum().item()),
                "radius_mean_degree": st["mean"], "nbr_sum": int(rg["nbr_count"].sum().item())}

    from concurrent.futures import ThreadPoolExecutor

    def lane_worker(lane, slides):
        # one host thread per lane: its slides run on its own stream and handle, so the H2D of one slide overlaps
        # the kernels (and the host reads of totals) of the other
        torch.cuda.set_device(local)
        out = {}
        with torch.cuda.stream(streams[lane]):
            for s in slides:
                out[s] = one_slide(s, engines[lane])
        streams[lane].synchronize()
        return out

    def run_mine():
        with ThreadPoolExecutor(max_workers=lanes) as ex:
            parts = list(ex.map(lane_worker, range(lanes), [mine[i::lanes] for i in range(lanes)]))
        merged = {}
        for d in parts:
            merged.update(d)
        return merged

    for lane in range(lanes):                                                     # warm-up (allocations, first launches)
        with torch.cuda.stream(streams[lane]):
            one_slide(mine[0], engines[lane])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    local_res = run_mine()
    torch.cuda.synchronize()
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, local_res)
        res = {}
        for d in gathered:
            res.update(d)
    else:
        res = local_res
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert sorted(res) == list(range(n_slides))
        chk = sum(v["knn_edges"] * 3 + v["radius_edges"] * 5 + v["knn_deg_sum"] + v["nbr_sum"] for v in res.values())
        area = sum(v["area_sum"] for v in res.values())
        sec = float(dt.item())
        print(f"cohort_check ok: world={world} slides={n_slides} x {n} nuclei: {sec * 1e3:.1f} ms "
              f"({n_slides * n / sec / 1e6:.0f} M nuclei/s from host tables, {sec / n_slides * world * 1e3:.2f} ms per slide per GPU)  "
              f"checksum={chk} area_sum={area:.3f}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
