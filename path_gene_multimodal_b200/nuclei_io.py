"""Nuclei-table I/O <-> SoA / CSR.

The reference saves the WSI nuclei table twice (aggregated_hovernet_run.py:398-402):
``<slide>_hovernet_nuclei_wsi.parquet`` and the CSV twin, both straight from the DataFrame whose
``centroid`` / ``bounding_box`` / ``polygon`` / ``wsi_polygon`` cells are Python lists.  In Parquet those
columns are ``list<double>``, ``list<int64>`` and ``list<list<double>>`` - i.e. offsets + flat values, exactly
the CSR / SoA layout the kernels take - so this module moves tables between files and device arrays without a
Python object per nucleus:

    read_nuclei_table(path)                      parquet | csv  -> pyarrow.Table (reference schema)
    table_to_soa(table)                          Table -> NucleiSoA (numpy views of the Arrow buffers)
    add_wsi_coords_to_table(table, tiles_df)     Arrow-native twin of add_wsi_coords_to_nuclei
    write_nuclei_table(table, parquet, csv)      the two files the reference writes
    process_nuclei_file(...)                     file -> file (steps 4-5 of run_hovernet_pipeline_on_wsi_tiles)

The numeric body runs in pg_map_morph_* like the DataFrame entry point (no CPU fallback).
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import pandas as pd

from . import _host

WSI_COLUMNS = ["tile_key", "tile_x", "tile_y", "centroid_x", "centroid_y", "wsi_centroid_x", "wsi_centroid_y",
               "bbox_xmin", "bbox_ymin", "bbox_xmax", "bbox_ymax", "wsi_bbox_xmin", "wsi_bbox_ymin",
               "wsi_bbox_xmax", "wsi_bbox_ymax", "wsi_polygon"]
LIST_COLUMNS = ("bounding_box", "centroid", "polygon", "wsi_polygon")


@dataclass
class NucleiSoA:
    """Structure-of-arrays view of a nuclei table (what the C ABI takes)."""
    n: int
    centroid: np.ndarray          # float64 [N,2]
    bbox: np.ndarray              # int64 [N,4]
    poly_off: np.ndarray          # int32 [N+1]
    poly_xy: np.ndarray           # float64 [M,2]
    poly_is_none: np.ndarray      # bool [N]
    type: np.ndarray | None       # int64 [N]


def _pa():
    import pyarrow as pa

    return pa


def _combine(col):
    pa = _pa()
    if isinstance(col, pa.ChunkedArray):
        return col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
    return col


def _list_offsets(arr) -> np.ndarray:
    """int64 offsets of a (large_)list array, relative to its own first value."""
    return np.asarray(arr.offsets.to_numpy(zero_copy_only=False), dtype=np.int64)


def fixed_list_column(col, width: int, dtype, name: str) -> np.ndarray:
    """``list<T>`` column whose every cell has ``width`` items -> dense [N, width] array (no Python loop)."""
    pa = _pa()
    arr = _combine(col)
    n = len(arr)
    if n == 0:
        return np.zeros((0, width), dtype=dtype)
    if arr.null_count:
        raise ValueError(f"{name}: null cells are not allowed")
    if pa.types.is_fixed_size_list(arr.type):
        if arr.type.list_size != width:
            raise ValueError(f"{name}: cells must have {width} items")
        flat = arr.flatten()
    else:
        off = _list_offsets(arr)
        if not np.all(np.diff(off) == width):
            raise ValueError(f"{name}: cells must have {width} items")
        flat = arr.values.slice(int(off[0]), n * width)
    return np.asarray(flat.to_numpy(zero_copy_only=False)).astype(dtype, copy=False).reshape(n, width)


def polygon_column_to_csr(col, dtype=np.float64):
    """``list<list<double>>`` column -> (poly_off int32 [N+1], poly_xy [M,2], is_none bool [N])."""
    arr = _combine(col)
    n = len(arr)
    if n == 0:
        return np.zeros(1, dtype=np.int32), np.zeros((0, 2), dtype=dtype), np.zeros(0, dtype=bool)
    pa = _pa()
    if pa.types.is_null(arr.type):  # every polygon is None
        return np.zeros(n + 1, dtype=np.int32), np.zeros((0, 2), dtype=dtype), np.ones(n, dtype=bool)
    is_none = np.asarray(arr.is_null().to_numpy(zero_copy_only=False), dtype=bool)
    off = _list_offsets(arr)
    inner = arr.values
    base_v, m = int(off[0]), int(off[-1] - off[0])
    if pa.types.is_fixed_size_list(inner.type):
        if inner.type.list_size != 2:
            raise ValueError("polygon vertices must be [x, y] pairs")
        flat = inner.flatten().slice(2 * base_v, 2 * m)
    else:
        ioff = _list_offsets(inner)
        if m and not np.all(np.diff(ioff[base_v: base_v + m + 1]) == 2):
            raise ValueError("polygon vertices must be [x, y] pairs")
        flat = inner.values.slice(int(ioff[base_v]), 2 * m)
    if m > _host.INT32_MAX:
        raise OverflowError("more than 2^31 polygon vertices")
    xy = np.asarray(flat.to_numpy(zero_copy_only=False)).astype(dtype, copy=False).reshape(m, 2)
    # a null cell may carry a non-empty slot in foreign files: its vertices stay in poly_xy but the row is empty
    rel = (off - base_v)
    if is_none.any() and np.any(np.diff(rel)[is_none] != 0):
        keep = np.repeat(~is_none, np.diff(rel))
        xy = xy[keep]
        cnt = np.where(is_none, 0, np.diff(rel))
        rel = np.concatenate([[0], np.cumsum(cnt)])
    return rel.astype(np.int32), np.ascontiguousarray(xy), is_none


def csr_to_polygon_column(poly_off, poly_xy, is_none=None):
    """CSR rings -> Arrow ``list<list<double>>`` (null where ``is_none``): the type pandas gives a column of
    list-of-[x, y] lists, so the Parquet file reads back like the reference's."""
    pa = _pa()
    xy = np.ascontiguousarray(poly_xy, dtype=np.float64).reshape(-1)
    m = xy.size // 2
    inner = pa.ListArray.from_arrays(pa.array(np.arange(0, 2 * m + 1, 2, dtype=np.int32)), pa.array(xy))
    off = pa.array(np.asarray(poly_off, dtype=np.int32))
    mask = None
    if is_none is not None and np.any(is_none):
        mask = pa.array(np.asarray(is_none, dtype=bool))
    return pa.ListArray.from_arrays(off, inner, mask=mask)


def _dense_to_list_column(a: np.ndarray, pa_type):
    pa = _pa()
    n, w = a.shape
    return pa.ListArray.from_arrays(pa.array(np.arange(0, n * w + 1, w, dtype=np.int32)),
                                    pa.array(np.ascontiguousarray(a).reshape(-1), type=pa_type))


def _parse_list_strings(series: pd.Series):
    """CSV twin: a list column written by DataFrame.to_csv holds ``repr(list)`` strings (valid JSON) or NaN for
    None.  One json.loads over the whole column instead of one literal_eval per cell."""
    vals = series.to_numpy(dtype=object)
    null = pd.isna(vals)
    doc = "[" + ",".join("null" if z else str(v) for v, z in zip(vals, null)) + "]"
    if "nan" in doc or "inf" in doc:  # repr(list) writes non-finite floats as nan / inf / -inf; JSON spells them differently
        doc = re.sub(r"(?<![\w.])(-?)inf(?![\w.])", r"\1Infinity", re.sub(r"(?<![\w.])nan(?![\w.])", "NaN", doc))
    return json.loads(doc)


def read_nuclei_table(path, columns=None):
    """Parquet (preferred) or the CSV twin -> pyarrow.Table with list columns restored."""
    pa = _pa()
    path = Path(path)
    if path.suffix.lower() in (".parquet", ".pq"):
        import pyarrow.parquet as pq

        return pq.read_table(path, columns=columns)
    df = pd.read_csv(path, usecols=columns)
    cols = {}
    for name in df.columns:
        if name in LIST_COLUMNS:
            py = _parse_list_strings(df[name])
            typ = {"bounding_box": pa.list_(pa.int64()), "centroid": pa.list_(pa.float64())}.get(
                name, pa.list_(pa.list_(pa.float64())))
            cols[name] = pa.array(py, type=typ)
        else:
            cols[name] = pa.array(df[name], from_pandas=True)
    return pa.table(cols)


def table_to_soa(table, polygon_col: str = "polygon") -> NucleiSoA:
    """Reference-schema table -> SoA / CSR arrays (views of the Arrow buffers where dtypes already match)."""
    n = table.num_rows
    cent = fixed_list_column(table["centroid"], 2, np.float64, "centroid")
    bbox = fixed_list_column(table["bounding_box"], 4, np.int64, "bounding_box")
    off, xy, is_none = polygon_column_to_csr(table[polygon_col])
    typ = None
    if "type" in table.column_names:
        typ = np.asarray(_combine(table["type"]).to_numpy(zero_copy_only=False)).astype(np.int64, copy=False)
    return NucleiSoA(n=n, centroid=cent, bbox=bbox, poly_off=off, poly_xy=xy, poly_is_none=is_none, type=typ)


def _stem_codes(col):
    """(codes int64 [N], stems object [n_unique]) of a string column: Path(p).stem once per distinct value."""
    arr = _combine(col)
    enc = arr.dictionary_encode() if not _pa().types.is_dictionary(arr.type) else arr
    if enc.null_count:
        raise ValueError("tile key column holds nulls")
    codes = np.asarray(enc.indices.to_numpy(zero_copy_only=False), dtype=np.int64)
    stems = np.array([Path(p).stem for p in enc.dictionary.to_pylist()], dtype=object)
    return codes, stems


def add_wsi_coords_to_table(nuc_table, tiles_df: pd.DataFrame, tile_key_col_nuc: str = "tile_path",
                            tile_key_col_tiles: str = "png_path", morphology: bool = False, device=None):
    """Arrow-native twin of ``add_wsi_coords_to_nuclei`` (aggregated_hovernet_run.py:263-336): same appended
    columns, order, dtypes and ``ValueError``, but the table never becomes Python objects."""
    from .nuclei_wsi import MORPH_COLUMNS, map_morph_arrays

    pa = _pa()
    n = nuc_table.num_rows
    # ---- :285-299 key join (strings, host): stems once per distinct path, first tile row per stem
    t_keys = np.array([Path(p).stem for p in tiles_df[tile_key_col_tiles]], dtype=object)
    first = {}
    for i, k in enumerate(t_keys):
        first.setdefault(k, i)
    codes, stems = _stem_codes(nuc_table[tile_key_col_nuc]) if n else (np.empty(0, np.int64), np.empty(0, object))
    row_of_stem = np.array([first.get(s, -1) for s in stems], dtype=np.int64)
    tile_row = row_of_stem[codes] if n else np.empty(0, dtype=np.int64)
    if n and (tile_row < 0).any():
        missing = pd.unique(stems[codes][tile_row < 0])
        raise ValueError(f"Some nuclei have tile_key with no matching tile coords: {missing}")
    tiles_x, tiles_y = tiles_df["x"].to_numpy(), tiles_df["y"].to_numpy()
    soa = table_to_soa(nuc_table)
    res = map_morph_arrays(soa.poly_off, soa.poly_xy, nuc_tile=tile_row.astype(np.int32),
                           tile_x=_host.as_int32(tiles_x, "tiles_df.x"), tile_y=_host.as_int32(tiles_y, "tiles_df.y"),
                           centroid=soa.centroid, bbox=_host.as_int32(soa.bbox, "bounding_box"),
                           write_polygons=True, device=device) if n else None
    wsi_c = res["wsi_centroid"] if n else np.zeros((0, 2))
    idt = np.result_type(soa.bbox.dtype, tiles_x.dtype)
    wsi_b = res["wsi_bbox"].astype(idt) if n else np.zeros((0, 4), dtype=idt)
    new = {
        "tile_key": pa.DictionaryArray.from_arrays(pa.array(codes.astype(np.int32)), pa.array(list(stems), type=pa.string())).cast(pa.string())
        if n else pa.array([], type=pa.string()),
        "tile_x": pa.array(tiles_x[tile_row]) if n else pa.array([], type=pa.int64()),
        "tile_y": pa.array(tiles_y[tile_row]) if n else pa.array([], type=pa.int64()),
        "centroid_x": pa.array(np.ascontiguousarray(soa.centroid[:, 0])),
        "centroid_y": pa.array(np.ascontiguousarray(soa.centroid[:, 1])),
        "wsi_centroid_x": pa.array(np.ascontiguousarray(wsi_c[:, 0])),
        "wsi_centroid_y": pa.array(np.ascontiguousarray(wsi_c[:, 1])),
    }
    for c, name in enumerate(["bbox_xmin", "bbox_ymin", "bbox_xmax", "bbox_ymax"]):
        new[name] = pa.array(np.ascontiguousarray(soa.bbox[:, c]))
    for c, name in enumerate(["wsi_bbox_xmin", "wsi_bbox_ymin", "wsi_bbox_xmax", "wsi_bbox_ymax"]):
        new[name] = pa.array(np.ascontiguousarray(wsi_b[:, c]))
    new["wsi_polygon"] = csr_to_polygon_column(soa.poly_off, res["wsi_poly_xy"] if n else np.zeros((0, 2)), soa.poly_is_none)
    if morphology:
        for name in MORPH_COLUMNS:
            new[name] = pa.array(res[name].astype(np.float64)) if n else pa.array([], type=pa.float64())
    out = nuc_table
    for name, col in new.items():
        if name in out.column_names:
            out = out.set_column(out.column_names.index(name), name, col)
        else:
            out = out.append_column(name, col)
    return out


def write_nuclei_table(table, parquet_path=None, csv_path=None) -> None:
    """The two files of aggregated_hovernet_run.py:398-402.  Parquet is written from Arrow directly; the CSV twin
    needs ``repr(list)`` cells, which only the pandas writer produces (slow path, one Python object per cell)."""
    if parquet_path is not None:
        import pyarrow.parquet as pq

        pq.write_table(table, str(parquet_path))
    if csv_path is not None:
        df = table.to_pandas()
        for name in LIST_COLUMNS:
            if name in df.columns:  # Arrow hands back ndarray cells; the reference's CSV shows nested lists
                col = table[name].to_pylist()
                df[name] = pd.Series(col, index=df.index, dtype=object)
        df.to_csv(str(csv_path), index=False)


def process_nuclei_file(nuclei_path, tiles, out_parquet=None, out_csv=None, tile_key_col_nuc: str = "tile_path",
                        tile_key_col_tiles: str = "png_path", morphology: bool = False, device=None):
    """Tile-local nuclei table file -> WSI nuclei table file(s): steps 4-5 of run_hovernet_pipeline_on_wsi_tiles
    (aggregated_hovernet_run.py:387-402).  ``tiles`` is the tile-annotation DataFrame or a CSV path."""
    tiles_df = tiles if isinstance(tiles, pd.DataFrame) else pd.read_csv(tiles)
    table = read_nuclei_table(nuclei_path)
    out = add_wsi_coords_to_table(table, tiles_df, tile_key_col_nuc, tile_key_col_tiles, morphology=morphology, device=device)
    write_nuclei_table(out, out_parquet, out_csv)
    return out
