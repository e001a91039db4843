"""Raster region properties of a HoverNeXt instance map (first half).

Reference call sites: ``regionprops(inst_map)`` at aggregated_hovernet_run.py:172-181 (per-instance
``bounding_box = [x_min, y_min, x_max, y_max]``) and ``regionprops_table(inst_map, properties=...)`` at
hovernet_tile_inference.ipynb:2415-2429 (cell 18).  One CUDA pass pair over the map (pg_raster_props) replaces
skimage's per-region Python loop, and one thread per instance traces and simplifies its contour polygon
(pg_instance_contours_*, :183-198); ``solidity`` comes from an exact convex-hull rasterisation (pg_raster_solidity).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import _host
from .engine import get_engine
from .polygon_morphology import derived_columns, zscore_columns


def raster_regionprops(inst_map, device=None) -> pd.DataFrame:
    """``regionprops_table`` columns for every label present (ascending, like skimage): ``label, area, perimeter,
    eccentricity, major_axis_length, minor_axis_length, orientation, solidity, centroid-0, centroid-1, bbox-0 .. bbox-3``."""
    m = np.asarray(inst_map)
    if m.ndim == 3:  # aggregated_hovernet_run.py:165-166
        m = m[0]
    if m.ndim != 2:
        raise ValueError("inst_map must be 2-D")
    eng = get_engine(device)
    n_labels = int(m.max()) if m.size else 0
    cols = ["label", "area", "perimeter", "eccentricity", "major_axis_length", "minor_axis_length", "orientation",
            "solidity", "centroid-0", "centroid-1", "bbox-0", "bbox-1", "bbox-2", "bbox-3"]
    if n_labels <= 0:
        return pd.DataFrame({c: np.zeros(0, dtype=np.int64 if c in ("label", "area") or c.startswith("bbox") else np.float64)
                             for c in cols})
    with torch.cuda.device(eng.device):
        d_m = _host.to_device(_host.as_int32(m, "inst_map"), np.int32, eng.device)
        res = _host.to_host_many(eng.raster_props(d_m, n_labels, solidity=True))
    keep = res["area"] > 0
    lab = np.nonzero(keep)[0] + 1
    df = pd.DataFrame({
        "label": lab.astype(np.int64),
        "area": res["area"][keep].astype(np.float64),   # skimage >= 0.20 reports area as float
        "perimeter": res["perimeter"][keep],
        "eccentricity": res["eccentricity"][keep],
        "major_axis_length": res["major_axis"][keep],
        "minor_axis_length": res["minor_axis"][keep],
        "orientation": res["orientation"][keep],
        "solidity": res["solidity"][keep],
        "centroid-0": res["centroid"][keep, 0], "centroid-1": res["centroid"][keep, 1],
    })
    for c in range(4):
        df[f"bbox-{c}"] = res["bbox"][keep, c].astype(np.int64)
    return df


def raster_morphology_table(inst_map, zscore: bool = False, device=None) -> pd.DataFrame:
    """``morph_df`` of cell 18 (ipynb:2415-2456) from the raster, as the reference computes it: ``inst_id, area,
    perimeter, eccentricity, solidity, major_axis_length, minor_axis_length, orientation`` + the derived
    ``perimeter_area, compactness, roundness, elongation`` (+ ``*_z`` of cell 21)."""
    df = raster_regionprops(inst_map, device=device)
    out = df[["label", "area", "perimeter", "eccentricity", "solidity", "major_axis_length", "minor_axis_length", "orientation"]].rename(
        columns={"label": "inst_id"})
    out = derived_columns(out)
    return zscore_columns(out) if zscore else out


def instance_bounding_boxes(inst_map, device=None) -> dict:
    """``bbox_dict`` of aggregated_hovernet_run.py:172-181: inst_id -> [x_min, y_min, x_max, y_max] (max exclusive)."""
    df = raster_regionprops(inst_map, device=device)
    return {int(l): [int(c0), int(r0), int(c1), int(r1)]
            for l, r0, c0, r1, c1 in zip(df["label"], df["bbox-0"], df["bbox-1"], df["bbox-2"], df["bbox-3"])}


def instance_polygons_csr(inst_map, tolerance: float = 0.5, device=None):
    """(labels int64 [K], poly_off int32 [K+1], poly_xy float64 [M,2]) - the polygon of every label present, as CSR:
    ``find_contours(inst_map == id, 0.5)`` -> longest contour -> (x, y) -> ``approximate_polygon(tolerance)`` of
    aggregated_hovernet_run.py:183-198, one CUDA thread per instance (pg_instance_contours_*). Rings are closed
    (first == last vertex) like skimage's; Douglas-Peucker ties are decided in exact arithmetic."""
    m = np.asarray(inst_map)
    if m.ndim == 3:
        m = m[0]
    if m.ndim != 2:
        raise ValueError("inst_map must be 2-D")
    eng = get_engine(device)
    n_labels = int(m.max()) if m.size else 0
    if n_labels <= 0:
        return np.zeros(0, dtype=np.int64), np.zeros(1, dtype=np.int32), np.zeros((0, 2))
    with torch.cuda.device(eng.device):
        d_m = _host.to_device(_host.as_int32(m, "inst_map"), np.int32, eng.device)
        rp = eng.raster_props(d_m, n_labels)
        res = eng.instance_contours(d_m, n_labels, rp["area"], rp["bbox"], tolerance)
        host = _host.to_host_many({"off": res["poly_off"], "xy": res["poly_xy"], "area": rp["area"]})
    nv = np.diff(host["off"].astype(np.int64))
    keep = (host["area"] > 0) & (nv > 0)
    lab = np.nonzero(keep)[0]
    xy_idx = np.concatenate([np.arange(host["off"][l], host["off"][l + 1]) for l in lab]) if len(lab) else np.zeros(0, dtype=np.int64)
    off = np.zeros(len(lab) + 1, dtype=np.int32)
    off[1:] = np.cumsum(nv[lab])
    return (lab + 1).astype(np.int64), off, np.ascontiguousarray(host["xy"][xy_idx.astype(np.int64)]).reshape(-1, 2)


def instance_polygons(inst_map, tolerance: float = 0.5, device=None) -> dict:
    """``poly_dict`` of aggregated_hovernet_run.py:183-198: inst_id -> list of [x, y]."""
    from .nuclei_wsi import csr_to_polygons

    labels, off, xy = instance_polygons_csr(inst_map, tolerance, device)
    rings = csr_to_polygons(off, xy)
    return {int(l): rings[i] for i, l in enumerate(labels)}


# PanNuke type ids of the HoverNeXt checkpoint the reference uses (aggregated_hovernet_run.py:76-82)
TYPE_NAMES = {1: "neoplastic", 2: "inflammatory", 3: "connective", 4: "dead", 5: "epithelial"}


def tile_nuclei_table(class_info: dict, inst_map, png_path, type_names: dict = TYPE_NAMES, device=None) -> pd.DataFrame:
    """Steps 1 and 3-5 of ``run_hovernet_on_tile`` (aggregated_hovernet_run.py:135-223) from HoverNeXt's two outputs:
    ``class_info`` (class_inst.json: ``{inst_id: [type, [0, cx, cy]]}``) and the instance map (pinst_pp).

    Returns the reference's ``final_df``: ``nuc_id, inst_id, type, type_name, bounding_box, centroid, polygon,
    tile_name, tile_path``; ``bounding_box`` / ``polygon`` are ``None`` for an id that is not in the map (``dict.get``
    semantics of :200-201). The per-instance masking loop becomes two CUDA passes (K12 + K13)."""
    import uuid
    from pathlib import Path

    rows = []
    for key, val in class_info.items():                       # :140-157
        _, cx, cy = val[1]
        rows.append({"inst_id": int(key), "type": int(val[0]), "centroid": [float(cx), float(cy)]})
    if not rows:
        return pd.DataFrame()
    df = pd.DataFrame(rows)
    bbox_dict = instance_bounding_boxes(inst_map, device=device)
    poly_dict = instance_polygons(inst_map, device=device)
    df["bounding_box"] = df["inst_id"].map(bbox_dict.get)     # :200-201
    df["polygon"] = df["inst_id"].map(poly_dict.get)
    df["type_name"] = df["type"].map(type_names)              # :204-205
    df["nuc_id"] = df["inst_id"].apply(lambda _: uuid.uuid4().hex)
    png_path = Path(png_path)
    df["tile_name"] = png_path.stem                            # :208-209
    df["tile_path"] = str(png_path)
    return df[["nuc_id", "inst_id", "type", "type_name", "bounding_box", "centroid", "polygon", "tile_name", "tile_path"]]
