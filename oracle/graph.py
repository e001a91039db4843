"""Oracle (test infrastructure): kNN / radius cell graphs, undirected union, composition, degree.

Follows /root/reference/hovernet_tile_inference.ipynb:
* cell 11 (ipynb:1815-1850)  ``KNN.from_array(coords, k)`` (libpysal: cKDTree.query(k+1) minus self),
  per-neighbour ``sqrt(dx*dx + dy*dy)``;
* cell 11 (ipynb:1865-1894)  undirected ``nx.Graph`` union with ``weight = min(dist)``;
* cell 12 (ipynb:1989-2002)  type-filtered sub-graph;
* cells 23-26 (ipynb:2963-3042)  ``cKDTree.query_ball_tree(tree, r)``, ``i < j`` edge list,
  ``edge_index`` / ``edge_attr``.
Composition / degree statistics are build-defined (README.md:127,136 names them only).

Canonical order (north_star): kNN rows ascending by ``(d^2, index)`` with self removed by index;
radius rows ascending by ``j``.  ``d^2 = fl(fl(dx*dx) + fl(dy*dy))`` in float64 - what scipy's
``sqeuclidean_distance_double`` evaluates for m = 2 on a non-FMA x86-64 build.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def _d2(a, b):
    dx = a[..., 0] - b[..., 0]
    dy = a[..., 1] - b[..., 1]
    return dx * dx + dy * dy


def knn_bruteforce(coords, k):
    """O(N^2) canonical kNN for small N: (idx int64 [N,k], dist float64 [N,k])."""
    c = np.asarray(coords, dtype=np.float64)
    n = len(c)
    if k >= n:
        raise ValueError("k must be smaller than the number of points")
    idx = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=np.float64)
    ar = np.arange(n)
    for i in range(n):
        d2 = _d2(c[i][None, :], c)
        keep = ar != i
        order = np.lexsort((ar[keep], d2[keep]))[:k]
        idx[i] = ar[keep][order]
        dist[i] = np.sqrt(d2[keep][order])
    return idx, dist


def knn(coords, k, workers=-1):
    """Canonical kNN through scipy's cKDTree (the reference's engine under libpysal).

    query(k + 1 + pad) -> drop self by index -> re-rank by (d^2, idx) -> rows whose k-th and
    (k+1)-th candidates tie (so the tree's arbitrary tie choice could matter) are redone by
    brute force over all points within that distance.
    """
    c = np.ascontiguousarray(coords, dtype=np.float64)
    n = len(c)
    if k >= n:
        raise ValueError("k must be smaller than the number of points")
    tree = cKDTree(c)
    kk = min(n, k + 2)
    _, nb = tree.query(c, k=kk, workers=workers)
    nb = nb.reshape(n, kk)
    ar = np.arange(n)
    d2 = _d2(c[:, None, :], c[nb])
    is_self = nb == ar[:, None]
    # make sure self is dropped even when duplicates pushed it out of the first slot / the list
    d2key = np.where(is_self, -1.0, d2)
    # two-key sort per row: by idx first, then stable by d2
    o1 = np.argsort(nb, axis=1, kind="stable")
    nb1 = np.take_along_axis(nb, o1, axis=1)
    dk1 = np.take_along_axis(d2key, o1, axis=1)
    o2 = np.argsort(dk1, axis=1, kind="stable")
    nb2 = np.take_along_axis(nb1, o2, axis=1)
    dk2 = np.take_along_axis(dk1, o2, axis=1)
    has_self = dk2[:, 0] < 0
    start = np.where(has_self, 1, 0)
    cols = start[:, None] + np.arange(k)[None, :]
    idx = np.take_along_axis(nb2, cols, axis=1)
    dd = np.take_along_axis(dk2, cols, axis=1)
    # rows needing an exact redo: self missing from the list (duplicates), or a tie at the k boundary
    nxt_col = np.minimum(start + k, kk - 1)
    nxt = np.take_along_axis(dk2, nxt_col[:, None], axis=1)[:, 0]
    boundary_tie = (start + k <= kk - 1) & (nxt == dd[:, -1])
    redo = np.nonzero(boundary_tie | ~has_self | (start + k > kk))[0]
    for i in redo:
        rad = np.sqrt(dd[i, -1]) * (1.0 + 1e-12) + 1e-300
        cand = np.array(tree.query_ball_point(c[i], rad), dtype=np.int64)
        cand = cand[cand != i]
        cd2 = _d2(c[i][None, :], c[cand])
        o = np.lexsort((cand, cd2))[:k]
        idx[i] = cand[o]
        dd[i] = cd2[o]
    return idx.astype(np.int64), np.sqrt(dd)


def undirected_union(knn_idx, knn_dist):
    """Cell-11 nx.Graph semantics over index arrays, vectorised.

    Returns (edges int64 [E,2] with i<j sorted by (i,j), weight float64 [E] = min over directions,
    row_ptr int64 [N+1], col int64 [2E] sorted per row, w float64 [2E]).
    """
    idx = np.asarray(knn_idx, dtype=np.int64)
    n, k = idx.shape
    src = np.repeat(np.arange(n, dtype=np.int64), k)
    dst = idx.reshape(-1)
    w = np.asarray(knn_dist, dtype=np.float64).reshape(-1)
    keep = src != dst
    src, dst, w = src[keep], dst[keep], w[keep]
    lo, hi = np.minimum(src, dst), np.maximum(src, dst)
    key = lo * n + hi
    order = np.lexsort((w, key))
    key, lo, hi, w = key[order], lo[order], hi[order], w[order]
    first = np.ones(len(key), dtype=bool)
    first[1:] = key[1:] != key[:-1]
    lo, hi, w = lo[first], hi[first], w[first]
    edges = np.stack([lo, hi], axis=1)
    row_ptr, col, ww = symmetric_csr(edges, w, n)
    return edges, w, row_ptr, col, ww


def undirected_union_networkx(knn_idx, knn_dist):
    """Literal restatement of the cell-11 loop (ipynb:1879-1894) with networkx; small N only."""
    import networkx as nx

    g = nx.Graph()
    n, k = np.asarray(knn_idx).shape
    g.add_nodes_from(range(n))
    for i in range(n):
        for s in range(k):
            j = int(knn_idx[i][s])
            if i == j:
                continue
            d = float(knn_dist[i][s])
            if g.has_edge(i, j):
                g.edges[i, j]["weight"] = min(g.edges[i, j]["weight"], d)
            else:
                g.add_edge(i, j, weight=d)
    return g


def symmetric_csr(edges, w, n):
    """i<j edge list -> symmetric CSR with rows sorted by column."""
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    src = np.concatenate([e[:, 0], e[:, 1]])
    dst = np.concatenate([e[:, 1], e[:, 0]])
    ww = np.concatenate([w, w])
    order = np.lexsort((dst, src))
    src, dst, ww = src[order], dst[order], ww[order]
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=row_ptr[1:])
    return row_ptr, dst, ww


def radius_graph(coords, r):
    """Cells 23-26: query_ball_tree + i<j filter + norm + float32 edge_attr.

    Returns dict(edges int64 [E,2], dist float64 [E], edge_index int64 [2,2E] (hstack layout, see
    SURVEY B-3), edge_attr float32 [2E,1], row_ptr, col, csr_dist).
    """
    c = np.ascontiguousarray(coords, dtype=np.float64)
    n = len(c)
    tree = cKDTree(c)
    pairs = tree.query_ball_tree(tree, r)
    counts = np.fromiter((len(p) for p in pairs), dtype=np.int64, count=n)
    src = np.repeat(np.arange(n, dtype=np.int64), counts)
    dst = np.fromiter((j for p in pairs for j in p), dtype=np.int64, count=int(counts.sum()))
    keep = src < dst
    e = np.stack([src[keep], dst[keep]], axis=1)
    order = np.lexsort((e[:, 1], e[:, 0]))
    e = e[order]
    return _pack_radius(c, e, n)


def radius_graph_notebook_loop(coords, r):
    """The literal cell-23 Python loop (ipynb:2969-2975); used for timing and small-N checks."""
    c = np.ascontiguousarray(coords, dtype=np.float64)
    tree = cKDTree(c)
    pairs = tree.query_ball_tree(tree, r)
    edges = []
    for i, neighs in enumerate(pairs):
        for j in neighs:
            if i != j and i < j:
                edges.append((i, j))
    return np.array(edges, dtype=np.int64).reshape(-1, 2)


def radius_graph_bruteforce(coords, r):
    """O(N^2) check: d^2 <= r*r in float64, no FMA."""
    c = np.asarray(coords, dtype=np.float64)
    n = len(c)
    r2 = np.float64(r) * np.float64(r)
    out = []
    for i in range(n):
        d2 = _d2(c[i][None, :], c)
        j = np.nonzero(d2 <= r2)[0]
        j = j[j > i]
        out.append(np.stack([np.full(len(j), i, dtype=np.int64), j], axis=1))
    e = np.concatenate(out) if out else np.zeros((0, 2), dtype=np.int64)
    return _pack_radius(c, e, n)


def _pack_radius(c, e, n):
    d = np.sqrt(_d2(c[e[:, 0]], c[e[:, 1]]))  # == np.linalg.norm(c[e0]-c[e1], axis=1) (ipynb:3041)
    edge_index = np.hstack([e.T, e[:, ::-1].T])
    edge_attr = np.concatenate([d[:, None], d[:, None]], axis=0).astype(np.float32)
    row_ptr, col, dd = symmetric_csr(e, d, n)
    return {"edges": e, "dist": d, "edge_index": edge_index, "edge_attr": edge_attr,
            "row_ptr": row_ptr, "col": col, "csr_dist": dd}


def composition(row_ptr, col, types, n_types=5):
    """nbr_count[i, t-1] = #{j in N(i): type[j] == t}, t = 1..T. int32 [N,T]."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n = len(row_ptr) - 1
    src = np.repeat(np.arange(n), np.diff(row_ptr))
    t = np.asarray(types, dtype=np.int64)[np.asarray(col, dtype=np.int64)]
    ok = (t >= 1) & (t <= n_types)
    out = np.zeros((n, n_types), dtype=np.int32)
    np.add.at(out, (src[ok], t[ok] - 1), 1)
    return out


def degree_stats(row_ptr):
    """degree int32 [N] and {min, max, sum, sumsq, mean, std(ddof=0), hist}."""
    deg = np.diff(np.asarray(row_ptr, dtype=np.int64))
    n = len(deg)
    if n == 0:
        return deg.astype(np.int32), {"min": 0, "max": 0, "sum": 0, "sumsq": 0, "mean": float("nan"),
                                      "std": float("nan"), "hist": np.zeros(1, dtype=np.int64)}
    stats = {
        "min": int(deg.min()), "max": int(deg.max()), "sum": int(deg.sum()),
        "sumsq": int((deg * deg).sum()), "mean": float(deg.mean()), "std": float(deg.std()),
        "hist": np.bincount(deg, minlength=int(deg.max()) + 1).astype(np.int64),
    }
    return deg.astype(np.int32), stats


def filter_types(edges, types, keep_types=(1, 2)):
    """Cell 12 (ipynb:1989-2002): keep nodes whose type is in keep_types, edges with both ends kept."""
    types = np.asarray(types)
    keep_node = np.isin(types, list(keep_types))
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    keep_edge = keep_node[e[:, 0]] & keep_node[e[:, 1]]
    return np.nonzero(keep_node)[0], e[keep_edge]
