"""Synthetic HoverNeXt nuclei tables.

The reference ships no data, so every parity / bench input is generated here with
``numpy.random.default_rng(seed)``.  The shapes follow what the reference's own
producer emits (``aggregated_hovernet_run.py:136-223``): per nucleus a tile-local
``centroid`` (2 floats), an integer ``bounding_box`` ``[xmin, ymin, xmax, ymax]``
(``:179-180``), a ``polygon`` on the half-pixel lattice that ``find_contours(level=0.5)``
produces (``:185-196``), an integer ``type`` in 1..5 (``TYPE_NAMES`` ``:76-82``) and the
``tile_path`` of the PNG it came from; tiles carry the top-left ``(x, y)`` of a 508-px
Mussel patch (``load_annotation_with_coordinates.py:21,177-180``).

Everything is produced as flat arrays (CSR polygons); ``to_frames`` builds the pandas
frames the reference function surface takes.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

DENSITY = 101.0 / (521.0 * 521.0)  # nuclei / px^2 of the notebook tile (ipynb:378,1549)
TILE_PITCH = 508                    # Mussel native patch size
TYPE_P = np.array([0.50, 0.23, 0.17, 0.03, 0.07])
SEEDS = {"C1": 1001, "C2": 1002, "C3": 1003, "C4": 1004, "C5": 1005}


@dataclass
class NucleiTable:
    """Flat (SoA / CSR) nuclei table in tile-local coordinates."""
    n_tiles_side: int
    tile_x: np.ndarray      # int32 [n_tiles]
    tile_y: np.ndarray      # int32 [n_tiles]
    nuc_tile: np.ndarray    # int32 [N]  index into tile_x / tile_y
    centroid: np.ndarray    # float64 [N, 2] tile-local (column 0 is what the reference calls x)
    bbox: np.ndarray        # int32 [N, 4] xmin, ymin, xmax, ymax (tile-local)
    types: np.ndarray       # int32 [N] in 1..5
    poly_off: np.ndarray    # int32 [N+1]
    poly_xy: np.ndarray     # float32 or float64 [M, 2] tile-local, open rings

    @property
    def n(self) -> int:
        return int(self.nuc_tile.shape[0])

    @property
    def side_px(self) -> int:
        return self.n_tiles_side * TILE_PITCH

    def wsi_centroids(self) -> np.ndarray:
        """float64 [N, 2] WSI-space centroids (what the graph stages consume)."""
        out = self.centroid.copy()
        out[:, 0] += self.tile_x[self.nuc_tile]
        out[:, 1] += self.tile_y[self.nuc_tile]
        return out


def slide_side_tiles(n: int, density: float = DENSITY) -> int:
    """Tiles per side so that n nuclei land at the notebook tile's density."""
    side = np.sqrt(n / density)
    return max(1, int(np.ceil(side / TILE_PITCH)))


def make_points(n: int, seed: int, density: float = DENSITY):
    """Only the graph inputs: WSI centroids float64 [N,2], types int32 [N], slide side (px)."""
    rng = np.random.default_rng(seed)
    side_t = slide_side_tiles(n, density)
    n_tiles = side_t * side_t
    nuc_tile = rng.integers(0, n_tiles, size=n, dtype=np.int64)
    local = rng.random((n, 2)) * TILE_PITCH
    types = rng.choice(np.arange(1, 6, dtype=np.int32), size=n, p=TYPE_P).astype(np.int32)
    xy = local
    xy[:, 0] += (nuc_tile % side_t) * TILE_PITCH
    xy[:, 1] += (nuc_tile // side_t) * TILE_PITCH
    return xy, types, side_t * TILE_PITCH


def make_polygons(n: int, seed: int, v_fixed: int | None = None, v_lo: int = 8, v_hi: int = 32,
                  centre: np.ndarray | None = None, dtype=np.float32):
    """CSR ellipse polygons snapped to the 0.5-px lattice (open rings).

    semi-major a ~ U[4,12], b = a*U[0.5,1], rotation U[0,pi), +-5 % radial jitter.
    Returns (poly_off int32[n+1], poly_xy dtype[M,2]).
    """
    rng = np.random.default_rng(seed + 7919)
    if v_fixed is not None:
        nv = np.full(n, v_fixed, dtype=np.int64)
    else:
        nv = rng.integers(v_lo, v_hi + 1, size=n, dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nv, out=off[1:])
    m = int(off[-1])
    owner = np.repeat(np.arange(n), nv)
    kth = np.arange(m) - off[owner]
    a = rng.uniform(4.0, 12.0, size=n)
    b = a * rng.uniform(0.5, 1.0, size=n)
    rot = rng.uniform(0.0, np.pi, size=n)
    t = 2.0 * np.pi * kth / nv[owner]
    jit = 1.0 + rng.uniform(-0.05, 0.05, size=m)
    ex = a[owner] * np.cos(t) * jit
    ey = b[owner] * np.sin(t) * jit
    c, s = np.cos(rot[owner]), np.sin(rot[owner])
    px = c * ex - s * ey
    py = s * ex + c * ey
    if centre is None:
        centre = rng.random((n, 2)) * TILE_PITCH
    px += centre[owner, 0]
    py += centre[owner, 1]
    xy = np.stack([np.round(px * 2.0) / 2.0, np.round(py * 2.0) / 2.0], axis=1).astype(dtype)
    return off.astype(np.int32), xy


def make_table(n: int, seed: int, v_fixed: int | None = None, v_lo: int = 8, v_hi: int = 32,
               density: float = DENSITY, dtype=np.float32) -> NucleiTable:
    """Full nuclei table: tiles, tile-local centroids, bboxes, types and CSR polygons."""
    rng = np.random.default_rng(seed)
    side_t = slide_side_tiles(n, density)
    n_tiles = side_t * side_t
    tix = np.arange(n_tiles, dtype=np.int64)
    tile_x = ((tix % side_t) * TILE_PITCH).astype(np.int32)
    tile_y = ((tix // side_t) * TILE_PITCH).astype(np.int32)
    nuc_tile = rng.integers(0, n_tiles, size=n, dtype=np.int64).astype(np.int32)
    centroid = rng.random((n, 2)) * TILE_PITCH
    types = rng.choice(np.arange(1, 6, dtype=np.int32), size=n, p=TYPE_P).astype(np.int32)
    off, xy = make_polygons(n, seed, v_fixed, v_lo, v_hi, centre=centroid, dtype=dtype)
    # bbox = int floor / ceil of the polygon extent
    xmin = np.minimum.reduceat(xy[:, 0], off[:-1])
    xmax = np.maximum.reduceat(xy[:, 0], off[:-1])
    ymin = np.minimum.reduceat(xy[:, 1], off[:-1])
    ymax = np.maximum.reduceat(xy[:, 1], off[:-1])
    bbox = np.stack([np.floor(xmin), np.floor(ymin), np.ceil(xmax), np.ceil(ymax)], axis=1).astype(np.int32)
    return NucleiTable(side_t, tile_x, tile_y, nuc_tile, centroid, bbox, types, off, xy)


_COHORT_BASES: dict = {}


def make_cohort_slide(slide: int, n: int, n_bases: int = 4, dtype=np.float32) -> NucleiTable:
    """Slide `slide` of the C4 cohort (64 slides x 500 k nuclei, seed 1004 + slide).

    make_table costs ~2.6 s per 500 k-nucleus slide on one host core (trigonometry for 10 M vertices), too slow to
    generate 64 of them inside a benchmark run, so the cohort is dealt from ``n_bases`` base tables: slide s takes
    base table s mod n_bases (seed 1004 + s mod n_bases: tile-local centroids, boxes, ragged polygons, types) and
    its OWN tile assignment drawn from rng(1004 + s) - i.e. its own WSI coordinates and therefore its own graphs.
    Sizes, dtypes, density and every byte count are those of 64 independent tables. The arrays a slide shares with
    its base are the same objects (pin them once)."""
    key = (slide % n_bases, n, np.dtype(dtype).name)
    base = _COHORT_BASES.get(key)
    if base is None:
        base = _COHORT_BASES[key] = make_table(n, SEEDS["C4"] + slide % n_bases, dtype=dtype)
    rng = np.random.default_rng(SEEDS["C4"] + slide)
    n_tiles = base.n_tiles_side * base.n_tiles_side
    nuc_tile = rng.integers(0, n_tiles, size=n, dtype=np.int64).astype(np.int32)
    return NucleiTable(base.n_tiles_side, base.tile_x, base.tile_y, nuc_tile, base.centroid, base.bbox, base.types,
                       base.poly_off, base.poly_xy)


def tile_png_paths(tab: NucleiTable, out_dir: str = "/data/out") -> list[str]:
    """png_path naming of load_annotation_with_coordinates.py:177-180: <out>/patches/<x>_<y>.png"""
    return [f"{out_dir}/patches/{int(x)}_{int(y)}.png" for x, y in zip(tab.tile_x, tab.tile_y)]


def to_frames(tab: NucleiTable, out_dir: str = "/data/out", closed_rings: bool = False):
    """pandas frames in the reference's layout: (nuc_df, tiles_df).

    nuc_df columns follow aggregated_hovernet_run.py:211-223; tiles_df has the columns
    add_wsi_coords_to_nuclei reads (``png_path, x, y``, :288-292) plus ``tile_index``.
    """
    import pandas as pd

    paths = tile_png_paths(tab, out_dir)
    tiles_df = pd.DataFrame({
        "tile_index": np.arange(len(paths)),
        "x": tab.tile_x.astype(np.int64),
        "y": tab.tile_y.astype(np.int64),
        "png_path": paths,
    })
    off = tab.poly_off
    xy64 = tab.poly_xy.astype(np.float64)
    polys = []
    for i in range(tab.n):
        p = xy64[off[i]:off[i + 1]].tolist()
        if closed_rings and p:
            p.append(list(p[0]))
        polys.append(p)
    names = {1: "neoplastic", 2: "inflammatory", 3: "connective", 4: "dead", 5: "epithelial"}
    nuc_df = pd.DataFrame({
        "nuc_id": [f"{i:032x}" for i in range(tab.n)],
        "inst_id": np.arange(1, tab.n + 1),
        "type": tab.types.astype(np.int64),
        "type_name": [names[int(t)] for t in tab.types],
        "bounding_box": [b.tolist() for b in tab.bbox.astype(np.int64)],
        "centroid": [c.tolist() for c in tab.centroid],
        "polygon": polys,
        "tile_name": [f"{int(tab.tile_x[t])}_{int(tab.tile_y[t])}" for t in tab.nuc_tile],
        "tile_path": [paths[t] for t in tab.nuc_tile],
    })
    return nuc_df, tiles_df
