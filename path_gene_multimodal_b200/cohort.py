"""Slide-parallel cohort runs (BASELINE config 4): every slide is the reference's one-slide job
(/root/reference/main.py:322-335) - its nuclei table goes from HOST arrays through the whole hot path on one GPU:
H2D of the table, tile->WSI map + polygon morphology, kNN k=8 + undirected union + i<j edges + composition, radius
graph r=50 px + composition + degree statistics, and a small per-slide summary back.  The graphs stay on the GPU
that built them.  ``lanes`` slides are in flight per GPU, each on its own stream / handle / host thread, so the
H2D of one slide overlaps the kernels and the host reads of the other.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _host
from .engine import Engine, default_knn_cell, get_engine, radius_cell

TABLE_FIELDS = ("poly_off", "poly_xy", "nuc_tile", "tile_x", "tile_y", "centroid", "bbox", "types")


def to_halfpx(poly_xy: np.ndarray) -> np.ndarray:
    """Tile-local contour vertices as int16 half-pixels (q = 2 * coordinate). skimage.measure.find_contours(mask, 0.5)
    (aggregated_hovernet_run.py:185-197) puts every vertex on the half-pixel lattice, so this is exact for the
    reference's polygons; anything else is refused rather than rounded."""
    q = np.asarray(poly_xy, dtype=np.float64) * 2.0
    r = np.rint(q)
    if not (np.array_equal(q, r) and (np.abs(r) <= 32767).all()):
        raise ValueError("to_halfpx: vertices are not on the half-pixel lattice of a tile (or exceed int16)")
    return r.astype(np.int16)


def pin_table(tab, cache: dict | None = None, staging: str = "float"):
    """Page-lock the arrays of a synth.NucleiTable in place (as a loader reading Parquet straight into pinned buffers
    would leave them); arrays shared between tables are pinned once (``cache`` maps id(array) -> pinned copy).
    ``staging="halfpx16"`` stores the polygon vertices as int16 half-pixels (to_halfpx): half the bytes of float32
    over the host link, widened on the device (Engine.widen_halfpx) to the very same float32 values."""
    if staging not in ("float", "halfpx16"):
        raise ValueError("staging must be 'float' or 'halfpx16'")
    cache = {} if cache is None else cache
    for name in TABLE_FIELDS:
        a = getattr(tab, name)
        key = (id(a), staging if name == "poly_xy" else "")
        hit = cache.get(key)
        if hit is None:
            src = to_halfpx(a) if (name == "poly_xy" and staging == "halfpx16" and a.dtype != np.int16) else a
            buf = _host.pinned_empty(src.shape, src.dtype)
            buf[...] = src
            hit = cache[key] = (buf, a)      # keep the source alive: id() stays unique
        setattr(tab, name, hit[0])
    return tab


def table_bytes(tab) -> int:
    return int(sum(getattr(tab, name).nbytes for name in TABLE_FIELDS))


def process_slide(eng: Engine, tab, k: int = 8, r: float = 50.0, n_types: int = 5) -> dict:
    """One slide, host table in, summary out (runs on the current stream of the engine's device)."""
    dev = eng.device
    n = tab.n
    side_px = float(tab.n_tiles_side * 508)
    t_off = _host.to_device(tab.poly_off, np.int32, dev)
    if tab.poly_xy.dtype == np.int16:      # half-pixel staging: 4 bytes per vertex over the link, widened here (exact)
        t_xy = eng.widen_halfpx(_host.to_device(tab.poly_xy, np.int16, dev))
    else:
        t_xy = _host.to_device(tab.poly_xy, np.float32 if tab.poly_xy.dtype == np.float32 else np.float64, dev)
    t_tile = _host.to_device(tab.nuc_tile, np.int32, dev)
    t_tx, t_ty = _host.to_device(tab.tile_x, np.int32, dev), _host.to_device(tab.tile_y, np.int32, dev)
    t_cen, t_bb = _host.to_device(tab.centroid, np.float64, dev), _host.to_device(tab.bbox, np.int32, dev)
    t_types = _host.to_device(tab.types, np.int32, dev)
    mm = eng.map_morph(t_off, t_xy, t_tile, t_tx, t_ty, t_cen, t_bb, write_polygons=True)
    wsi = mm["wsi_centroid"]
    bnd = (0.0, 0.0, side_px, side_px)
    eng.grid_build(wsi, t_types, None, default_knn_cell(n, side_px ** 2, k), bnd)
    kn = eng.knn(k, dist_dtype=torch.float32)
    # union outputs sized by their bounds, radius outputs by the capacity a previous slide of this engine needed:
    # nothing is read back until the summary, so the slide is ONE enqueue + ONE synchronising read
    un = eng.knn_union(kn["knn_idx"], kn["dist32"], types=t_types, n_types=n_types, symmetric_dist=True, presized=True)
    eng.grid_build(wsi, t_types, None, radius_cell(r), bnd)
    cap = getattr(eng, "_cohort_radius_cap", None)
    if cap is None:
        rg = eng.radius_graph(r, upper=True, n_types=n_types, want_dist32=True, want_edges=True)
        cap = int(rg["total"])
    else:
        rg = eng.radius_graph(r, upper=True, n_types=n_types, want_dist32=True, want_edges=True, capacity=cap)

    def edge_hash(e, valid):
        # sum over the valid prefix (a device scalar) of i * 1000003 + j, in wrapping int64
        keep = torch.arange(e.shape[0], device=dev) < valid
        return torch.where(keep, e[:, 0] * 1000003 + e[:, 1], torch.zeros((), dtype=torch.int64, device=dev)).sum()

    n_knn, n_rad = un["up_ptr"][-1].long(), rg["row_ptr"][-1].long()
    sums = torch.stack([mm["area"].double().sum().reshape(1).view(torch.int64)[0], un["degree"].sum(), rg["nbr_count"].sum(),
                        edge_hash(un["edges"], n_knn), edge_hash(rg["edges"], n_rad), n_knn, n_rad,
                        rg["stats"][1], rg["stats"][3]]).tolist()
    if sums[6] > cap:   # the capacity hint was too small for this slide: once more, exactly
        eng.lib.pg_check_overflow(eng._h)
        eng._cohort_radius_cap = None
        return process_slide(eng, tab, k, r, n_types)
    eng._cohort_radius_cap = max(int(sums[6] * 1.25) + 1024, cap if cap < 2 * sums[6] else 0)
    return {"knn_edges": int(sums[5]), "radius_edges": int(sums[6]),
            "area_sum": float(np.int64(sums[0]).view(np.float64)), "knn_deg_sum": int(sums[1]), "nbr_sum": int(sums[2]),
            "knn_edge_hash": int(sums[3]), "radius_edge_hash": int(sums[4]),
            "radius_mean_degree": (sums[7] / sums[8]) if sums[8] else float("nan")}


class CohortRunner:
    """``lanes`` engines + streams on one device; ``run(slides, get_table)`` processes the slides and returns
    ({slide: summary}, device milliseconds between the first enqueue and the last completion)."""

    def __init__(self, device: int, lanes: int = 2):
        self.device = int(device)
        self.dev = torch.device("cuda", self.device)
        self.lanes = max(1, int(lanes))
        self.engines = [get_engine(self.device)] + [Engine(self.device) for _ in range(self.lanes - 1)]
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.lanes)]

    def close(self):
        for e in self.engines[1:]:
            e.close()

    def _lane(self, lane: int, slides, get_table, start_event):
        torch.cuda.set_device(self.device)
        out = {}
        with torch.cuda.stream(self.streams[lane]):
            self.streams[lane].wait_event(start_event)
            for s in slides:
                out[s] = process_slide(self.engines[lane], get_table(s))
        return out

    def run(self, slides, get_table):
        torch.cuda.set_device(self.device)
        main = torch.cuda.current_stream(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        slides = list(slides)
        with ThreadPoolExecutor(max_workers=self.lanes) as ex:
            parts = list(ex.map(lambda lane: self._lane(lane, slides[lane::self.lanes], get_table, e0), range(self.lanes)))
        for st in self.streams:
            main.wait_stream(st)
        e1.record(main)
        e1.synchronize()
        merged = {}
        for d in parts:
            merged.update(d)
        return merged, e0.elapsed_time(e1)


def cohort_checksum(results: dict) -> dict:
    """Digest of per-slide summaries that does not depend on the number of GPUs / lanes (slides in id order)."""
    vals = [results[s] for s in sorted(results)]
    mask = (1 << 63) - 1
    return {"edges": int(sum(v["knn_edges"] * 3 + v["radius_edges"] * 5 + v["knn_deg_sum"] + v["nbr_sum"] for v in vals)),
            "edge_hash": int(sum(v["knn_edge_hash"] + v["radius_edge_hash"] for v in vals) & mask),
            "area_sum": float(sum(v["area_sum"] for v in vals))}
