"""Export adapters into the object types the notebook hands on: ``nx.Graph`` (cell 11) - the only place in the
package that imports networkx, and only to CONSTRUCT the object from arrays the GPU path has already produced (no
graph algorithm of networkx runs; tests/test_host_cpu.py pins that)."""
from __future__ import annotations

import numpy as np


def to_networkx(graph: dict, df=None, node_ids=None):
    """The notebook's ``G_knn`` (ipynb:1865-1894) from a graph dict with ``edges`` int [E,2] and ``weight`` (or
    ``dist``): an ``nx.Graph`` whose nodes are keyed by ``df["nuc_id"]`` (or ``node_ids``, or the row number) and
    carry ``type_id, type_name, color, pos`` when ``df`` has them; edges carry ``weight``.  networkx (a dependency
    of the reference) is imported here only; the graph itself was built on the GPU."""
    import networkx as nx

    edges = np.asarray(graph["edges"])
    w = graph.get("weight", graph.get("dist"))
    n = int(len(df)) if df is not None else (len(node_ids) if node_ids is not None else int(edges.max()) + 1 if len(edges) else 0)
    if node_ids is None:
        node_ids = df["nuc_id"].tolist() if df is not None and "nuc_id" in df.columns else list(range(n))
    node_ids = list(node_ids)
    G = nx.Graph()
    pos = graph.get("pos")
    for i, nid in enumerate(node_ids):
        attrs = {}
        if df is not None:
            for name in ("type_id", "type_name", "color"):
                if name in df.columns:
                    attrs[name] = df[name].iloc[i]
        if pos is not None:
            attrs["pos"] = (float(pos[i][0]), float(pos[i][1]))
        G.add_node(nid, **attrs)
    ids = np.asarray(node_ids, dtype=object)
    if w is None:
        G.add_edges_from(zip(ids[edges[:, 0]], ids[edges[:, 1]]))
    else:
        G.add_weighted_edges_from(zip(ids[edges[:, 0]], ids[edges[:, 1]], np.asarray(w, dtype=np.float64).tolist()))
    return G
