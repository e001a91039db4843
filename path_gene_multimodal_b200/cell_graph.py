"""Spatial cell graphs over nucleus centroids, in the notebook's vocabulary.

Mirrors /root/reference/hovernet_tile_inference.ipynb:
* ``build_knn_graph``      cell 11 (ipynb:1766-1951): ``knn_neighbors``, ``knn_neighbor_distances``, and the
  undirected ``nx.Graph`` union (edges ``i<j``, ``weight = min(dist)``);
* ``build_radius_graph``   cells 23-26 (ipynb:2963-3042): ``edges`` int64 [E,2] (``i<j``), ``edge_index``,
  ``edge_attr`` float32 [2E,1];
* ``filter_graph_by_type`` cell 12 (ipynb:1989-2002);
* ``neighbour_type_composition`` / ``degree_stats``: named in README.md:127,136 only; defined in SURVEY A.5.

Host arrays in, host arrays out; everything in between is libpathgraph (uniform-grid binning, grid
queries, two-pass CSR) on the GPU.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _host, _lib
from .engine import default_knn_cell, get_engine, radius_cell

TYPE_NAMES = {1: "neoplastic", 2: "inflammatory", 3: "connective", 4: "dead", 5: "epithelial"}       # ipynb:1774-1780
CLASS_COLORS = {1: "tab:red", 2: "tab:green", 3: "tab:blue", 4: "tab:gray", 5: "tab:orange"}           # ipynb:1782-1788


def _prep(coords, types, eng):
    c = np.asarray(coords, dtype=np.float64)
    if c.ndim != 2 or c.shape[1] != 2:
        raise ValueError("coords must have shape (N, 2)")
    if c.shape[0] > _host.INT32_MAX:
        raise OverflowError("more than 2^31 points")
    # NaN / inf (cKDTree raises on them) are detected on the device by the histogram kernel and reported at the
    # first synchronisation - no extra pass over the coordinates on the host
    d_xy = _host.to_device(c, np.float64, eng.device)
    d_t = None
    if types is not None:
        t = np.asarray(types)
        if t.shape[0] != c.shape[0]:
            raise ValueError("types must have one entry per point")
        d_t = _host.to_device(_host.as_int32(t, "types"), np.int32, eng.device)
    return c, d_xy, d_t


def _bounds(c):
    if c.shape[0] == 0:
        return (0.0, 0.0, 1.0, 1.0)
    return (float(c[:, 0].min()), float(c[:, 1].min()), float(c[:, 0].max()), float(c[:, 1].max()))


def build_knn_graph(coords, k: int = 5, types=None, n_types: int = 5, undirected: bool = True,
                    bounds=None, cell_size: float | None = None, device=None, neighbor_coords: bool = False) -> dict:
    """kNN graph of cell 11.

    Returns ``knn_neighbors`` int64 [N,k] (ascending ``(d^2, index)``, self excluded by index),
    ``knn_neighbor_distances`` float64 [N,k] and, with ``undirected``: ``edges`` int64 [E,2] (``i<j``,
    sorted), ``weight`` float64 [E], symmetric CSR ``row_ptr`` / ``col`` / ``csr_weight``, ``degree`` int32 [N],
    ``nbr_count`` int32 [N,n_types] (when ``types`` is given) and ``degree_stats``.
    ``neighbor_coords`` adds ``knn_neighbor_coords`` float64 [N,k,2] (the (x, y) of every list entry, ipynb:1838-1840).
    Raises ``ValueError`` when ``k >= N`` (as cKDTree-backed KNN does for k+1 > N).
    """
    eng = get_engine(device)
    c, d_xy, d_t = _prep(coords, types, eng)
    n = c.shape[0]
    if not (1 <= k):
        raise ValueError("k must be >= 1")
    if k >= n:
        raise ValueError(f"k={k} must be smaller than the number of points ({n})")
    with torch.cuda.device(eng.device):
        b = bounds if bounds is not None else _bounds(c)
        if not np.isfinite(b).all():
            raise ValueError("coords must be finite")
        if cell_size is None:
            cell_size = default_knn_cell(n, max(b[2] - b[0], 1e-9) * max(b[3] - b[1], 1e-9), k)
        eng.grid_build(d_xy, d_t, None, cell_size, b)
        kn = eng.knn(k, dist_dtype=torch.float64)
        out = {"knn_neighbors": _host.to_host(kn["knn_idx"]).astype(np.int64),
               "knn_neighbor_distances": _host.to_host(kn["dist"])}
        if neighbor_coords:
            out["knn_neighbor_coords"] = _host.to_host(eng.knn_neighbor_coords(kn["knn_idx"], d_xy))
        eng.grid_check()
        if undirected:
            # union + i<j edges + composition + degree statistics: one fused pass chain (pg_knn_union_*)
            u = eng.knn_union(kn["knn_idx"], kn["dist"], types=d_t, n_types=n_types, symmetric_dist=True)
            host = _host.to_host_many({"edges": u["edges"], "weight": u["edge_w"], "row_ptr": u["row_ptr"], "col": u["col"],
                                       "csr_weight": u["w"], "degree": u["degree"], "stats": u["stats"], "hist": u["hist"],
                                       "nbr_count": u["nbr_count"]})
            out.update({
                "edges": host["edges"], "weight": host["weight"],
                "row_ptr": host["row_ptr"].astype(np.int64), "col": host["col"].astype(np.int64),
                "csr_weight": host["csr_weight"], "degree": host["degree"],
                "degree_stats": eng.decode_stats(host["stats"], host["hist"]),
            })
            if d_t is not None:
                out["nbr_count"] = host["nbr_count"]
    return out


def build_radius_graph(coords, r: float = 40.0, types=None, n_types: int = 5, mpp: float | None = None,
                       symmetric_csr: bool = False, bounds=None, device=None, outputs: str = "notebook",
                       count_dtype=np.int32) -> dict:
    """Radius graph of cells 23-26: ``d(i, j) <= r`` (inclusive, like query_ball_tree), ``i < j``.

    ``mpp`` scales pixel coordinates to micrometres first (``x_um = x_px * mpp``, ipynb:2046).
    ``outputs="notebook"`` (default) returns ``edges`` int64 [E,2] sorted by (i, j); ``edge_index`` int64 [2,2E] =
    hstack(edges.T, edges[:, ::-1].T) (the reference's vstack yields [4,E]; see SURVEY B-3); ``edge_attr`` float32
    [2E,1]; ``dist`` float32 [E]; ``degree`` int32 [N]; ``nbr_count`` int32 [N,n_types] (with ``types``);
    ``degree_stats``; and with ``symmetric_csr`` also ``row_ptr`` / ``col`` / ``csr_dist``.
    ``outputs="compact"`` returns the same graph without the redundant int64 tensors: ``edges`` int32 [E,2],
    ``dist`` float32 [E], ``degree``, ``nbr_count``, ``degree_stats`` - a third of the bytes that cross PCIe (the
    notebook tensors are ``graph_features.edge_index_from_edges(edges, dist)`` away, on whichever device consumes
    them), and the whole call is one enqueue + two synchronising copies once the handle has seen a slide of this kind.
    ``count_dtype`` (compact only): ``np.uint8`` / ``np.uint16`` narrow ``degree`` and ``nbr_count`` on the device
    before the copy (a quarter / half of their bytes); ``OverflowError`` if a count does not fit.
    """
    if np.dtype(count_dtype) not in (np.dtype(np.int32), np.dtype(np.uint8), np.dtype(np.uint16)):
        raise ValueError("count_dtype must be int32, uint8 or uint16")
    if outputs != "compact" and np.dtype(count_dtype) != np.dtype(np.int32):
        raise ValueError("count_dtype needs outputs='compact'")
    if outputs not in ("notebook", "compact"):
        raise ValueError("outputs must be 'notebook' or 'compact'")
    eng = get_engine(device)
    c = np.asarray(coords, dtype=np.float64)
    if mpp is not None:
        c = c * float(mpp)
    c, d_xy, d_t = _prep(c, types, eng)
    if not (r >= 0 and np.isfinite(r)):
        raise ValueError("r must be finite and >= 0")
    with torch.cuda.device(eng.device):
        eng.grid_build(d_xy, d_t, None, radius_cell(r), bounds)  # bounds=None: min/max reduced on the device
        if outputs == "compact":
            out = _radius_compact(eng, c.shape[0], r, n_types, d_t is not None, np.dtype(count_dtype))
        else:
            g = eng.radius_graph(r, upper=True, n_types=n_types, compose=True, want_dist32=False, want_edge_index=True)
            e = int(g["total"])
            host = _host.to_host_many({"edge_index": g["edge_index"], "edge_attr": g["edge_attr"], "degree": g["degree"],
                                       "nbr_count": g["nbr_count"] if d_t is not None else None,
                                       "stats": g["stats"], "hist": g["hist"]})
            out = {
                "edges": host["edge_index"][:, :e].T,      # view: rows (i, j), i < j, sorted by (i, j)
                "dist": host["edge_attr"][:e, 0],
                "edge_index": host["edge_index"],
                "edge_attr": host["edge_attr"],
                "degree": host["degree"],
                "degree_stats": eng.decode_stats(host["stats"], host["hist"]),
            }
            if d_t is not None:
                out["nbr_count"] = host["nbr_count"]
        if symmetric_csr:
            s = eng.radius_graph(r, upper=False, compose=False, stats=False, want_dist32=True)
            out["row_ptr"] = _host.to_host(s["row_ptr"]).astype(np.int64)
            out["col"] = _host.to_host(s["col"]).astype(np.int64)
            out["csr_dist"] = _host.to_host(s["dist32"])
    return out


def _radius_compact(eng, n: int, r: float, n_types: int, has_types: bool, count_dtype=np.dtype(np.int32)) -> dict:
    """The compact output set. With a capacity hint from an earlier slide (edges per nucleus seen on this handle)
    the build is pg_radius_graph - outputs given up front, nothing read back in between - followed by two batches of
    device-to-host copies: the per-nucleus arrays with the edge count, then exactly E edges (the capacity's head-room
    stays on the device). Without a hint, or when the hint proves too small, the exact count -> total -> fill sequence
    runs (and leaves a hint behind)."""
    hint = getattr(eng, "_radius_edges_per_point", None)
    for attempt in range(2):
        if hint is not None and attempt == 0 and n > 0:
            cap = int(hint * 1.15 * n) + 4096
            prev = getattr(eng, "_radius_cap", 0)
            if cap <= prev < 2 * cap:
                cap = prev  # same buffers as the last slide of this kind
            eng._radius_cap = cap
            g = eng.radius_graph(r, upper=True, n_types=n_types, compose=True, want_dist32=True, want_edges32=True,
                                 capacity=cap, out=getattr(eng, "_radius_compact_out", None))
            eng._radius_compact_out = {k: v for k, v in g.items() if k in ("col", "dist32", "edges32")}  # reused scratch
        else:
            g = eng.radius_graph(r, upper=True, n_types=n_types, compose=True, want_dist32=True, want_edges32=True)
            cap = int(g["total"])
        deg, nbr = g["degree"], (g["nbr_count"] if has_types else None)
        if count_dtype != np.dtype(np.int32):
            tdt = torch.uint8 if count_dtype == np.dtype(np.uint8) else torch.int16
            deg = eng.narrow_counts(deg, tdt)
            nbr = eng.narrow_counts(nbr, tdt) if nbr is not None else None
        # two batches of copies: the per-nucleus arrays and the edge count first, then exactly E edges - the capacity's
        # 15 % head-room never crosses the link (the second synchronisation costs less than copying it)
        host = _host.to_host_many({"degree": deg, "nbr_count": nbr, "stats": g["stats"], "hist": g["hist"],
                                   "row_end": g["row_ptr"][-1:]})
        e = int(host["row_end"][0])
        if e <= cap:
            # the edge copies are waited for by the flag read that follows them on the stream (pg_check_overflow /
            # pg_grid_check synchronise): one synchronisation for the copies, the overflow flag and the input flag
            host.update(_host.to_host_many({"edges": g["edges32"][:e], "dist": g["dist32"][:e]}, sync=False))
            if count_dtype != np.dtype(np.int32):
                rc = eng.lib.pg_check_overflow(eng._h)  # (also clears the flag; a non-finite input is reported first)
                if rc == _lib.PG_ERR_CAPACITY:
                    raise OverflowError(f"a degree / neighbour-type count does not fit {count_dtype}")
                eng._check(rc)
                if count_dtype == np.dtype(np.uint16):
                    host["degree"] = host["degree"].view(np.uint16)
                    if host.get("nbr_count") is not None:
                        host["nbr_count"] = host["nbr_count"].view(np.uint16)
            break
        eng.lib.pg_check_overflow(eng._h)  # clears the overflow flag the undersized pass has raised
        hint = None
    if n > 0:
        eng._radius_edges_per_point = max(e / n, 1e-3)
    if count_dtype == np.dtype(np.int32):
        eng.grid_check()  # synchronises (the edge copies above) and reports a non-finite input
    out = {"edges": host["edges"], "dist": host["dist"], "degree": host["degree"],
           "degree_stats": eng.decode_stats(host["stats"], host["hist"])}
    if has_types:
        out["nbr_count"] = host["nbr_count"]
    return out


def neighbour_type_composition(row_ptr, col, types, n_types: int = 5, device=None) -> np.ndarray:
    """``nbr_count[i, t-1] = #{j in N(i): type[j] == t}`` for t = 1..n_types over any CSR. int32 [N,T]."""
    eng = get_engine(device)
    with torch.cuda.device(eng.device):
        d_rp = _host.to_device(_host.as_int32(row_ptr, "row_ptr"), np.int32, eng.device)
        d_col = _host.to_device(_host.as_int32(col, "col"), np.int32, eng.device)
        d_t = _host.to_device(_host.as_int32(types, "types"), np.int32, eng.device)
        res = eng.compose_degree(d_rp, d_col, d_t, n_types)
        return _host.to_host(res["nbr_count"])


def degree_stats(row_ptr, hist_len: int = 64, device=None) -> dict:
    """``degree`` int32 [N] and ``min, max, sum, sumsq, mean, std`` (ddof=0, ipynb:2903 convention), ``hist``."""
    eng = get_engine(device)
    with torch.cuda.device(eng.device):
        d_rp = _host.to_device(_host.as_int32(row_ptr, "row_ptr"), np.int32, eng.device)
        res = eng.compose_degree(d_rp, None, None, 1, hist_len=hist_len, compose=False)
        out = eng.decode_stats(res["stats"], res["hist"])
        out["degree"] = _host.to_host(res["degree"])
        return out


def filter_graph_by_type(edges, types, keep_types=(1, 2)):
    """Cell 12 (ipynb:1989-2002): nodes whose type is kept, and the edges with both ends kept.

    Index bookkeeping on the host (two boolean masks); returns (kept node ids int64, edges int64 [E',2])."""
    t = np.asarray(types)
    keep = np.isin(t, list(keep_types))
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    return np.nonzero(keep)[0], e[keep[e[:, 0]] & keep[e[:, 1]]]


def clustering_coefficients(row_ptr, col, device=None) -> dict:
    """``triangles`` int32 [N] and ``clustering`` float64 [N] (= networkx.clustering) of a symmetric CSR with
    ascending rows (README.md:136 "clustering"); ``degree_centrality`` = degree / (N - 1) (networkx's definition)."""
    eng = get_engine(device)
    with torch.cuda.device(eng.device):
        d_rp = _host.to_device(_host.as_int32(row_ptr, "row_ptr"), np.int32, eng.device)
        d_col = _host.to_device(_host.as_int32(col, "col"), np.int32, eng.device)
        res = eng.clustering(d_rp, d_col)
        out = _host.to_host_many(res)
    n = len(out["clustering"])
    deg = np.diff(np.asarray(row_ptr, dtype=np.int64))
    out["degree_centrality"] = deg / (n - 1.0) if n > 1 else np.ones(n)
    return out


def type_interaction_matrix(types, nbr_count, n_types: int = 5, device=None) -> np.ndarray:
    """int64 [T,T]: number of directed edges from a node of type a+1 to a neighbour of type b+1 (README.md:133
    "cell-cell interaction patterns"); symmetric for an undirected graph, diagonal counts each edge twice."""
    eng = get_engine(device)
    with torch.cuda.device(eng.device):
        d_t = _host.to_device(_host.as_int32(types, "types"), np.int32, eng.device)
        d_c = _host.to_device(_host.as_int32(nbr_count, "nbr_count"), np.int32, eng.device)
        return _host.to_host(eng.type_interactions(d_t, d_c, n_types))


def knn_graph_frame(final_df, k: int = 5, centroid_col: str = "centroid", centroid_order: str = "yx",
                    type_col: str | None = None, device=None):
    """Cell 11 at DataFrame level (ipynb:1766-1951): returns ``(df, graph)``.

    ``df`` is a copy of ``final_df`` with the columns the notebook copies back (:1945-1951): ``type_id``, ``color``,
    ``knn_neighbors`` (list of k row positions), ``knn_neighbor_coords`` (list of (x, y) tuples),
    ``knn_neighbor_points`` (shapely Points when shapely is importable, else the same tuples) and
    ``knn_neighbor_distances`` (list of k floats).  ``centroid`` cells are [y, x] as HoverNeXt stores them (the
    notebook builds ``Point(c[1], c[0])``, :1799); pass ``centroid_order="xy"`` for already swapped data.
    ``graph`` is the dict of ``build_knn_graph`` (edges, weights, degree, composition);
    ``interop.to_networkx(graph, df)`` makes the notebook's ``G_knn``."""
    import pandas as pd

    df = final_df.copy()
    c = np.array(df[centroid_col].tolist(), dtype=np.float64).reshape(-1, 2)
    coords = np.ascontiguousarray(c[:, ::-1]) if centroid_order == "yx" else c
    if "type_id" not in df.columns:                                  # ipynb:1807-1813
        if type_col is None and "type_name" in df.columns:
            name_to_id = {v: kk for kk, v in TYPE_NAMES.items()}
            df["type_id"] = df["type_name"].map(name_to_id)
        elif (type_col or "type") in df.columns:
            df["type_id"] = df[type_col or "type"]
        else:
            raise ValueError("No type_name or type column found to map cell types.")
    tid = pd.to_numeric(df["type_id"], errors="coerce")
    types = tid.fillna(0).astype(np.int64).to_numpy()
    g = build_knn_graph(coords, k=k, types=types, device=device, neighbor_coords=True)
    df["knn_neighbors"] = pd.Series(g["knn_neighbors"].tolist(), index=df.index, dtype=object)
    nc = g["knn_neighbor_coords"]
    as_tuples = [[(float(x), float(y)) for x, y in row] for row in nc.tolist()]
    df["knn_neighbor_coords"] = pd.Series(as_tuples, index=df.index, dtype=object)
    try:
        from shapely.geometry import Point  # the notebook's type; absent here -> the coordinate tuples stand in

        df["knn_neighbor_points"] = pd.Series([[Point(x, y) for x, y in row] for row in as_tuples], index=df.index, dtype=object)
    except ImportError:
        df["knn_neighbor_points"] = df["knn_neighbor_coords"]
    df["knn_neighbor_distances"] = pd.Series(g["knn_neighbor_distances"].tolist(), index=df.index, dtype=object)
    df["color"] = [CLASS_COLORS.get(int(t), "black") if np.isfinite(t) else "black" for t in tid.to_numpy(dtype=np.float64)]
    g["pos"] = coords
    return df, g
