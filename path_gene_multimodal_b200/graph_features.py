"""Node / edge feature assembly for the GNN input.

Reference: hovernet_tile_inference.ipynb:2903 (cell 21, z-scores), :2950 (cell 23, type one-hot and the
feature-column order), :3021-3042 (cells 25-26, ``edge_index`` / ``edge_attr``), :3068 (cell 27, PyG ``Data``).
``x`` itself is never defined in the notebook (its printed shape is [101, 15] = 5 one-hot + 10 z-scores): here it
is ``final_df[onehot_cols + morph_z_cols]`` as float32, built by one CUDA kernel pair (pg_node_features).
The notebook's ``onehot_cols`` also catches ``type_name`` and duplicates from re-running the cell (its printed
``feat_cols``); those artefacts are not reproduced.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import _host
from .cell_graph import build_radius_graph
from .engine import get_engine
from .polygon_morphology import CONT_COLS


def node_feature_matrix(df: pd.DataFrame, cont_cols=CONT_COLS, type_col: str = "type", device=None) -> dict:
    """``x`` float32 [N, n_types_present + n_cont] and its column names.

    One-hot columns follow ``pd.get_dummies(df[type_col], prefix="type")`` (one per distinct value, ascending);
    z-score columns follow cell 21 (``cont_cols`` that exist in ``df``, NaN-skipping mean / std(ddof=0),
    0.0 for a constant or empty column)."""
    eng = get_engine(device)
    cols = [c for c in cont_cols if c in df.columns]
    n = len(df)
    feat = np.ascontiguousarray(np.stack([df[c].to_numpy(dtype=np.float64) for c in cols])) if cols else None
    values = None
    names = []
    t = None
    if type_col is not None and type_col in df.columns:
        t = _host.as_int32(df[type_col].to_numpy(), type_col)
        values = np.unique(t)
        names += [f"type_{v}" for v in values]
    names += [c + "_z" for c in cols]
    with torch.cuda.device(eng.device):
        d_feat = _host.to_device(feat, np.float64, eng.device) if feat is not None else None
        d_t = _host.to_device(t, np.int32, eng.device) if t is not None else None
        d_v = _host.to_device(values.astype(np.int32), np.int32, eng.device) if values is not None else None
        if d_feat is None and d_t is None:
            return {"x": np.zeros((n, 0), dtype=np.float32), "columns": [], "mean": np.zeros(0), "std": np.zeros(0)}
        res = eng.node_features(d_feat, d_t, d_v)
        host = _host.to_host_many({"x": res["x"], "stats": res["stats"]})
    stats = host["stats"].reshape(-1, 2)
    return {"x": host["x"], "columns": names, "mean": stats[:, 0].copy(), "std": stats[:, 1].copy()}


def assemble_graph_data(df: pd.DataFrame, r: float = 40.0, coord_cols=("x_um", "y_um"), cont_cols=CONT_COLS,
                        type_col: str = "type", n_types: int = 5, device=None) -> dict:
    """Cells 23-27 in one call: ``x``, ``edge_index`` int64 [2,2E], ``edge_attr`` float32 [2E,1], ``edges``,
    ``pos`` - the arguments of ``torch_geometric.data.Data`` (see ``to_pyg``)."""
    coords = df[list(coord_cols)].to_numpy(dtype=np.float64)
    types = df[type_col].to_numpy() if type_col in df.columns else None
    g = build_radius_graph(coords, r=r, types=types, n_types=n_types, device=device)
    nf = node_feature_matrix(df, cont_cols=cont_cols, type_col=type_col, device=device)
    return {"x": nf["x"], "feat_cols": nf["columns"], "edge_index": g["edge_index"], "edge_attr": g["edge_attr"],
            "edges": g["edges"], "pos": coords, "degree": g["degree"], "nbr_count": g.get("nbr_count")}


def edge_index_from_edges(edges, dist):
    """The notebook's packed tensors from a compact edge list (``build_radius_graph(outputs="compact")``):
    ``edge_index`` int64 [2,2E] = hstack(edges.T, edges[:, ::-1].T) (ipynb:3021, SURVEY B-3) and ``edge_attr``
    float32 [2E,1] = concat(d, d) (ipynb:3041-3042). numpy in, numpy out; torch tensors (any device) in, torch out -
    so the widening to int64 happens where the graph is consumed, not before it crosses PCIe."""
    if isinstance(edges, torch.Tensor):
        e = edges.to(torch.int64)
        ei = torch.cat([e.t(), e.flip(1).t()], dim=1).contiguous()
        d = dist.to(torch.float32).reshape(-1, 1)
        return ei, torch.cat([d, d], dim=0)
    e = np.asarray(edges).astype(np.int64)
    d = np.asarray(dist, dtype=np.float32).reshape(-1, 1)
    return np.ascontiguousarray(np.hstack([e.T, e[:, ::-1].T])), np.concatenate([d, d], axis=0)


def to_pyg(data: dict):
    """``torch_geometric.data.Data(x=, edge_index=, edge_attr=)`` of cell 27 (needs torch_geometric)."""
    try:
        from torch_geometric.data import Data
    except ImportError as e:  # not a fallback: PyG is the consumer, not part of the path
        raise ImportError("to_pyg needs torch_geometric") from e
    return Data(x=torch.from_numpy(np.ascontiguousarray(data["x"])),
                edge_index=torch.from_numpy(np.ascontiguousarray(data["edge_index"])),
                edge_attr=torch.from_numpy(np.ascontiguousarray(data["edge_attr"])),
                pos=torch.from_numpy(np.ascontiguousarray(data["pos"])))
