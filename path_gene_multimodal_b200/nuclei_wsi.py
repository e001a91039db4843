"""Tile -> WSI coordinate map of the HoverNeXt nuclei table (drop-in for the reference function).

``add_wsi_coords_to_nuclei`` keeps the signature, the appended columns, their order and dtypes and
the ``ValueError`` of /root/reference/aggregated_hovernet_run.py:263-336.  The string-key join
(:285-299) stays on the host; the numeric body (:302-334) and, optionally, the polygon morphology
run in one fused CUDA kernel (pg_map_morph_*, csrc/pg_morph.cu).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pandas as pd
import torch

from . import _host
from .engine import get_engine

MORPH_COLUMNS = ["area", "perimeter", "eccentricity", "circularity"]


def polygons_to_csr(polygons, dtype=np.float64):
    """Column of rings (list of [x, y] lists, or None) -> (poly_off int32 [N+1], poly_xy [M,2], is_none bool [N]).

    Goes through Arrow: ``list<list<double>>`` is already CSR (outer offsets + flat interleaved
    values), so no Python loop touches the vertices.
    """
    import pyarrow as pa

    arr = None
    if isinstance(polygons, pd.Series):
        if isinstance(polygons.dtype, pd.ArrowDtype):   # an Arrow-backed column (wsi_polygon_as="arrow", read_parquet with
            chunks = polygons.array._pa_array            # dtype_backend="pyarrow"): its buffers ARE the CSR
            arr = (chunks.combine_chunks() if isinstance(chunks, pa.ChunkedArray) else chunks).cast(pa.list_(pa.list_(pa.float64())))
        else:
            polygons = polygons.to_numpy()
    n = len(polygons)
    if n == 0:
        return np.zeros(1, dtype=np.int32), np.zeros((0, 2), dtype=dtype), np.zeros(0, dtype=bool)
    if arr is None:
        try:
            arr = pa.array(polygons, type=pa.list_(pa.list_(pa.float64())), from_pandas=True)
        except (pa.ArrowInvalid, pa.ArrowTypeError, pa.ArrowNotImplementedError):
            # cells that are numpy arrays (what pd.read_parquet returns for list<list<double>>, including the
            # reference's own *_nuclei_wsi.parquet): normalise ring by ring
            return _polygons_to_csr_loop(polygons, dtype)
    is_none = np.asarray(arr.is_null().to_numpy(zero_copy_only=False), dtype=bool)
    off = np.asarray(arr.offsets.to_numpy(), dtype=np.int64)
    inner = arr.values  # list<double>, one entry per vertex
    ioff = np.asarray(inner.offsets.to_numpy(), dtype=np.int64)
    if len(ioff) > 1 and not np.all(np.diff(ioff) == 2):
        raise ValueError("polygon vertices must be [x, y] pairs")
    flat = np.asarray(inner.values.to_numpy(zero_copy_only=False), dtype=np.float64)
    base_v = int(off[0])
    m = int(off[-1]) - base_v
    xy = flat[int(ioff[base_v]): int(ioff[base_v]) + 2 * m].reshape(m, 2)
    off = off - base_v
    if m > _host.INT32_MAX:
        raise OverflowError("more than 2^31 polygon vertices")
    return off.astype(np.int32), np.ascontiguousarray(xy, dtype=dtype), is_none


def _polygons_to_csr_loop(polygons, dtype):
    n = len(polygons)
    is_none = np.zeros(n, dtype=bool)
    off = np.zeros(n + 1, dtype=np.int64)
    rings = []
    for i, p in enumerate(polygons):
        if p is None or (isinstance(p, float) and np.isnan(p)):
            is_none[i] = True
            off[i + 1] = off[i]
            continue
        a = np.asarray([np.asarray(v, dtype=np.float64) for v in p], dtype=np.float64) if len(p) else np.zeros((0, 2))
        if a.ndim != 2 or (a.size and a.shape[1] != 2):
            raise ValueError("polygon vertices must be [x, y] pairs")
        rings.append(a.reshape(-1, 2))
        off[i + 1] = off[i] + a.shape[0]
    if off[-1] > _host.INT32_MAX:
        raise OverflowError("more than 2^31 polygon vertices")
    xy = np.concatenate(rings, axis=0) if rings else np.zeros((0, 2))
    return off.astype(np.int32), np.ascontiguousarray(xy, dtype=dtype), is_none


def csr_to_polygons(poly_off, poly_xy, is_none=None):
    """Inverse of polygons_to_csr: object array of list-of-[x, y] lists (None where is_none).

    The reference's column format (one Python list per vertex) is what costs here, not the conversion: ndarray.tolist
    builds every [x, y] once in C and the rings are slices of that list (6x faster than Arrow's to_pylist)."""
    n = len(poly_off) - 1
    out = np.empty(n, dtype=object)
    if n == 0:
        return out
    flat = np.ascontiguousarray(poly_xy, dtype=np.float64).reshape(-1, 2).tolist()
    off = np.asarray(poly_off).tolist()
    for i in range(n):
        out[i] = flat[off[i]:off[i + 1]]
    if is_none is not None:
        for i in np.nonzero(is_none)[0]:
            out[i] = None
    return out


def csr_to_arrow_series(poly_off, poly_xy, is_none=None, index=None) -> pd.Series:
    """CSR rings -> pandas Series with ArrowDtype(list<list<double>>) over the same buffers (no Python object per
    vertex); None where ``is_none``."""
    import pyarrow as pa

    xy = np.ascontiguousarray(poly_xy, dtype=np.float64).reshape(-1, 2)
    m = xy.shape[0]
    inner = pa.ListArray.from_arrays(pa.array(np.arange(0, 2 * m + 1, 2, dtype=np.int32)), pa.array(xy.reshape(-1)))
    mask = pa.array(np.asarray(is_none, dtype=bool)) if is_none is not None and np.any(is_none) else None
    outer = pa.ListArray.from_arrays(pa.array(np.asarray(poly_off, dtype=np.int32)), inner, mask=mask)
    return pd.Series(pd.arrays.ArrowExtensionArray(outer), index=index)


def map_morph_arrays(poly_off, poly_xy, nuc_tile=None, tile_x=None, tile_y=None, centroid=None, bbox=None,
                     write_polygons=True, extra=False, device=None) -> dict:
    """Array-level entry: numpy (host) in, numpy out, one fused kernel in between.

    poly_xy float32 or float64 [M,2] (the dtype chooses pg_map_morph_f32 / _f64).  Returns dict with
    wsi_poly_xy, wsi_centroid (float64 [N,2]), wsi_bbox (int32 [N,4]), area, perimeter, eccentricity,
    circularity (float32 [N]) and with ``extra`` major_axis, minor_axis, centroid_x, centroid_y, poly_bbox.
    """
    eng = get_engine(device)
    dev = eng.device
    vt = np.float32 if np.asarray(poly_xy).dtype == np.float32 else np.float64
    d_off = _host.to_device(poly_off, np.int32, dev)
    d_xy = _host.to_device(np.asarray(poly_xy).reshape(-1, 2), vt, dev)
    d_tile = _host.to_device(nuc_tile, np.int32, dev) if nuc_tile is not None else None
    d_tx = _host.to_device(tile_x, np.int32, dev) if tile_x is not None else None
    d_ty = _host.to_device(tile_y, np.int32, dev) if tile_y is not None else None
    d_c = _host.to_device(np.asarray(centroid).reshape(-1, 2), np.float64, dev) if centroid is not None else None
    d_b = _host.to_device(np.asarray(bbox).reshape(-1, 4), np.int32, dev) if bbox is not None else None
    with torch.cuda.device(dev):
        res = eng.map_morph(d_off, d_xy, d_tile, d_tx, d_ty, d_c, d_b, write_polygons=write_polygons, extra=extra)
        return {k: _host.to_host(v) for k, v in res.items()}


def _stems(values) -> tuple[np.ndarray, np.ndarray]:
    """(codes int64 [N], stems object [n_unique]): Path(p).stem evaluated once per distinct path."""
    codes, uniques = pd.factorize(np.asarray(values, dtype=object), use_na_sentinel=False)
    stems = np.array([Path(p).stem for p in uniques], dtype=object)
    return codes, stems


def add_wsi_coords_to_nuclei(
    nuc_df: pd.DataFrame,
    tiles_df: pd.DataFrame,
    tile_key_col_nuc: str = "tile_path",
    tile_key_col_tiles: str = "png_path",
    morphology: bool = False,
    device=None,
    centroid_order: str = "xy",
    wsi_polygon_as: str = "lists",
) -> pd.DataFrame:
    """Shift tile-local centroid / bounding_box / polygon by the tile's top-left (x, y).

    Same contract as the reference (aggregated_hovernet_run.py:263-336): returns a copy of
    ``nuc_df`` with ``tile_key, tile_x, tile_y, centroid_x, centroid_y, wsi_centroid_x,
    wsi_centroid_y, bbox_*, wsi_bbox_*, wsi_polygon`` appended in that order; inputs are not
    modified; unmatched tile keys raise ``ValueError``.  Like the reference, ``centroid[0]`` is
    treated as x (SURVEY B-1): HoverNeXt stores centroids as (y, x) (hovernet_plotting.py:65-66), so the
    reference's ``wsi_centroid_x`` is really tile_x + y.  ``centroid_order="yx"`` is the opt-in fix: ``centroid[1]``
    is x, ``centroid[0]`` is y, and the four centroid columns come out in true WSI axes (polygons and boxes are
    (x, y) either way).  ``morphology=True`` additionally appends ``area, perimeter, eccentricity, circularity``
    from the same kernel launch.  Tile offsets must be integral pixel values (int or float dtype).
    ``wsi_polygon_as="arrow"`` returns ``wsi_polygon`` as an Arrow-backed ``list<list<double>>`` column over the
    kernel's output buffers instead of Python lists (same values; ``.tolist()`` / element access give the lists): the
    reference's cell format costs three Python objects per vertex, which is what bounds this function, not the GPU.
    """
    if wsi_polygon_as not in ("lists", "arrow"):
        raise ValueError("wsi_polygon_as must be 'lists' (the reference's cells) or 'arrow'")
    if centroid_order not in ("xy", "yx"):
        raise ValueError("centroid_order must be 'xy' (the reference's reading) or 'yx' (HoverNeXt's storage order)")
    out = nuc_df.copy()
    n = len(out)
    # ---- :285-299 key join on the host (strings); stem once per distinct path
    t_codes, t_stems = _stems(tiles_df[tile_key_col_tiles])
    tile_key_per_row = t_stems[t_codes] if len(t_codes) else np.empty(0, dtype=object)
    first = pd.Series(np.arange(len(tile_key_per_row))).groupby(tile_key_per_row, sort=False).first() \
        if len(tile_key_per_row) else pd.Series(dtype=np.int64)
    lut = {k: int(v) for k, v in first.items()}  # stem -> first tiles row with that stem (:288-292)
    n_codes, n_stems = _stems(out[tile_key_col_nuc]) if n else (np.empty(0, dtype=np.int64), np.empty(0, dtype=object))
    row_of_stem = np.array([lut.get(s, -1) for s in n_stems], dtype=np.int64)
    if n and (row_of_stem[n_codes] < 0).any():
        missing = pd.unique(n_stems[n_codes][row_of_stem[n_codes] < 0])
        raise ValueError(f"Some nuclei have tile_key with no matching tile coords: {missing}")
    out["tile_key"] = pd.Series(n_stems[n_codes] if n else [], index=out.index, dtype=tiles_df[tile_key_col_tiles].dtype
                                if str(tiles_df[tile_key_col_tiles].dtype) == "str" else object)
    tiles_x = tiles_df["x"].to_numpy()
    tiles_y = tiles_df["y"].to_numpy()
    tile_row = row_of_stem[n_codes] if n else np.empty(0, dtype=np.int64)
    out["tile_x"] = tiles_x[tile_row] if n else np.empty(0, dtype=tiles_x.dtype)
    out["tile_y"] = tiles_y[tile_row] if n else np.empty(0, dtype=tiles_y.dtype)

    # ---- host lists -> SoA / CSR
    cent = np.array(out["centroid"].tolist(), dtype=np.float64).reshape(-1, 2) if n else np.zeros((0, 2))
    if centroid_order == "yx":
        cent = np.ascontiguousarray(cent[:, ::-1])
    bb_raw = np.array(out["bounding_box"].tolist()).reshape(-1, 4) if n else np.zeros((0, 4), dtype=np.int64)
    bb = _host.as_int32(bb_raw, "bounding_box")
    poly_off, poly_xy, is_none = polygons_to_csr(out["polygon"])
    res = map_morph_arrays(poly_off, poly_xy, nuc_tile=tile_row.astype(np.int32),
                           tile_x=_host.as_int32(tiles_x, "tiles_df.x"), tile_y=_host.as_int32(tiles_y, "tiles_df.y"),
                           centroid=cent, bbox=bb, write_polygons=True, device=device) if n else None

    # ---- :302-334 columns, in the reference's order and dtypes
    out["centroid_x"] = cent[:, 0]
    out["centroid_y"] = cent[:, 1]
    wsi_c = res["wsi_centroid"] if n else np.zeros((0, 2))
    out["wsi_centroid_x"] = wsi_c[:, 0]
    out["wsi_centroid_y"] = wsi_c[:, 1]
    for c, name in enumerate(["bbox_xmin", "bbox_ymin", "bbox_xmax", "bbox_ymax"]):
        out[name] = bb_raw[:, c]
    wsi_b = res["wsi_bbox"] if n else np.zeros((0, 4), dtype=np.int64)
    for c, name in enumerate(["wsi_bbox_xmin", "wsi_bbox_ymin", "wsi_bbox_xmax", "wsi_bbox_ymax"]):
        # bbox + tile_x for the x columns, bbox + tile_y for the y columns: numpy's result dtype of each sum (:316-319)
        out[name] = wsi_b[:, c].astype(np.result_type(bb_raw.dtype, (tiles_x if c % 2 == 0 else tiles_y).dtype))
    if wsi_polygon_as == "arrow":
        out["wsi_polygon"] = csr_to_arrow_series(poly_off, res["wsi_poly_xy"] if n else np.zeros((0, 2)), is_none, out.index)
    else:
        out["wsi_polygon"] = pd.Series(csr_to_polygons(poly_off, res["wsi_poly_xy"], is_none) if n else [],
                                       index=out.index, dtype=object)
    if morphology:
        for name in MORPH_COLUMNS:
            out[name] = res[name].astype(np.float64) if n else np.zeros(0)
    return out
