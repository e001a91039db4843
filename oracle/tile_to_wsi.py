"""Oracle (test infrastructure): tile -> WSI coordinate map.

Restates ``add_wsi_coords_to_nuclei`` (/root/reference/aggregated_hovernet_run.py:263-336)
and, where /root/reference is mounted (build container only), loads the reference's own
function by AST extraction so the restatement can be pinned against it.
"""
from __future__ import annotations

import ast
from pathlib import Path, PurePosixPath

import numpy as np
import pandas as pd

REFERENCE_FILE = Path("/root/reference/aggregated_hovernet_run.py")

NEW_COLUMNS = [
    "tile_key", "tile_x", "tile_y",
    "centroid_x", "centroid_y", "wsi_centroid_x", "wsi_centroid_y",
    "bbox_xmin", "bbox_ymin", "bbox_xmax", "bbox_ymax",
    "wsi_bbox_xmin", "wsi_bbox_ymin", "wsi_bbox_xmax", "wsi_bbox_ymax",
    "wsi_polygon",
]


def load_reference_function():
    """The reference's own add_wsi_coords_to_nuclei, compiled from the read-only file.

    The module cannot be imported (top-level zarr / skimage imports, :8-10), so only the one
    FunctionDef is compiled with {np, pd, Path} in scope.  Returns None when the reference
    tree is not mounted (e.g. on the GPU box).
    """
    if not REFERENCE_FILE.is_file():
        return None
    tree = ast.parse(REFERENCE_FILE.read_text())
    fn = next(n for n in tree.body
              if isinstance(n, ast.FunctionDef) and n.name == "add_wsi_coords_to_nuclei")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np, "pd": pd, "Path": Path}
    exec(compile(mod, str(REFERENCE_FILE), "exec"), ns)  # noqa: S102 - read-only reference source
    return ns["add_wsi_coords_to_nuclei"]


def add_wsi_coords_to_nuclei_oracle(nuc_df, tiles_df, tile_key_col_nuc="tile_path",
                                    tile_key_col_tiles="png_path"):
    """Plain-Python restatement (per-row loops, float64 / int64), column order as :273-334."""
    out = nuc_df.copy()
    # :285-286  key = filename stem on both sides
    tile_keys = [PurePosixPath(str(p)).stem for p in tiles_df[tile_key_col_tiles]]
    nuc_keys = [PurePosixPath(str(p)).stem for p in out[tile_key_col_nuc]]
    # :288-292  first occurrence per key wins
    lut = {}
    for key, x, y in zip(tile_keys, tiles_df["x"], tiles_df["y"]):
        if key not in lut:
            lut[key] = (x, y)
    missing = [k for k in dict.fromkeys(nuc_keys) if k not in lut]
    if missing:  # :297-299
        raise ValueError(f"Some nuclei have tile_key with no matching tile coords: {np.array(missing, dtype=object)}")
    out["tile_key"] = nuc_keys
    tx = np.array([lut[k][0] for k in nuc_keys], dtype=np.asarray(tiles_df["x"]).dtype).reshape(-1)
    ty = np.array([lut[k][1] for k in nuc_keys], dtype=np.asarray(tiles_df["y"]).dtype).reshape(-1)
    out["tile_x"] = tx
    out["tile_y"] = ty
    cent = np.array([list(c) for c in out["centroid"]], dtype=np.float64).reshape(-1, 2)  # :302-304
    out["centroid_x"] = cent[:, 0]
    out["centroid_y"] = cent[:, 1]
    out["wsi_centroid_x"] = tx + cent[:, 0]  # :306-307
    out["wsi_centroid_y"] = ty + cent[:, 1]
    bb = np.array([list(b) for b in out["bounding_box"]]).reshape(-1, 4)  # :310-314
    for c, name in enumerate(["bbox_xmin", "bbox_ymin", "bbox_xmax", "bbox_ymax"]):
        out[name] = bb[:, c]
    out["wsi_bbox_xmin"] = bb[:, 0] + tx  # :316-319
    out["wsi_bbox_ymin"] = bb[:, 1] + ty
    out["wsi_bbox_xmax"] = bb[:, 2] + tx
    out["wsi_bbox_ymax"] = bb[:, 3] + ty
    polys = []
    for poly, dx, dy in zip(out["polygon"], tx, ty):  # :322-334
        if poly is None:
            polys.append(None)
        else:
            polys.append([[float(x) + float(dx), float(y) + float(dy)] for x, y in poly])
    out["wsi_polygon"] = polys
    return out


def map_arrays(tile_x, tile_y, nuc_tile, centroid, bbox, poly_off, poly_xy):
    """Array-level oracle of the same map on the SoA / CSR layout the C ABI takes."""
    tx = np.asarray(tile_x, dtype=np.int64)[nuc_tile]
    ty = np.asarray(tile_y, dtype=np.int64)[nuc_tile]
    wsi_c = np.asarray(centroid, dtype=np.float64).copy()
    wsi_c[:, 0] += tx
    wsi_c[:, 1] += ty
    wsi_b = np.asarray(bbox, dtype=np.int64).copy()
    wsi_b[:, 0] += tx
    wsi_b[:, 2] += tx
    wsi_b[:, 1] += ty
    wsi_b[:, 3] += ty
    nv = np.diff(np.asarray(poly_off, dtype=np.int64))
    owner = np.repeat(np.arange(len(nv)), nv)
    wsi_p = np.asarray(poly_xy, dtype=np.float64).copy()
    wsi_p[:, 0] += tx[owner]
    wsi_p[:, 1] += ty[owner]
    return wsi_c, wsi_b, wsi_p
