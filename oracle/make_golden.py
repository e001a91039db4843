"""Generate the committed fixtures under tests/golden/ (run in the build container only).

    python -m oracle.make_golden

* add_wsi_ref.json      - inputs and outputs of the REFERENCE's own add_wsi_coords_to_nuclei
                          (AST-extracted from /root/reference/aggregated_hovernet_run.py:263-336,
                          executed unmodified) on a seeded 96-nucleus table with None polygons,
                          closed rings, duplicate tile stems and float / int mixes.
* graph_small.npz       - a seeded 600-point set (uniform + lattice + duplicates) with the outputs of
                          scipy.spatial.cKDTree.query_ball_tree (the notebook's exact call, ipynb:2964-2967),
                          cKDTree.query (what KNN.from_array wraps) re-ranked canonically, and brute force.
* notebook_known_answers.json - numbers printed in the stored notebook outputs.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pandas as pd
import scipy
from scipy.spatial import cKDTree

from oracle import graph, tile_to_wsi
from path_gene_multimodal_b200 import synth

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def make_add_wsi():
    ref = tile_to_wsi.load_reference_function()
    if ref is None:
        raise SystemExit("/root/reference is not mounted; fixtures can only be regenerated in the build container")
    tab = synth.make_table(96, seed=4242, v_lo=3, v_hi=40, dtype=np.float64)
    nuc, tiles = synth.to_frames(tab, out_dir="/scratch/run 1/out", closed_rings=False)
    # quirks the reference handles: None polygons, explicitly closed rings, a duplicated tile stem in
    # another directory (first occurrence must win), unused tiles, non-lattice float vertices
    polys = nuc["polygon"].tolist()
    polys[3] = None
    polys[10] = None
    polys[5] = polys[5] + [polys[5][0]]
    polys[7] = [[x + 0.123456789, y - 0.987654321] for x, y in polys[7]]
    nuc["polygon"] = polys
    dup = tiles.iloc[[0]].copy()
    dup["png_path"] = "/elsewhere/patches/" + Path(tiles.iloc[0]["png_path"]).name
    dup["x"] = 999999
    dup["y"] = 888888
    tiles = pd.concat([tiles, dup], ignore_index=True)
    out = ref(nuc, tiles, tile_key_col_nuc="tile_path", tile_key_col_tiles="png_path")
    payload = {
        "generator": "oracle/make_golden.py: reference add_wsi_coords_to_nuclei, aggregated_hovernet_run.py:263-336",
        "pandas": pd.__version__, "numpy": np.__version__,
        "nuc_df": json.loads(nuc.to_json(orient="split", double_precision=15)),
        "tiles_df": json.loads(tiles.to_json(orient="split", double_precision=15)),
        "out_columns": list(out.columns),
        "out_dtypes": {c: str(t) for c, t in out.dtypes.items()},
        "out": json.loads(out.to_json(orient="split", double_precision=15)),
    }
    # exact float round trip: store hex for the float columns
    payload["out_hex"] = {
        c: [float(v).hex() for v in out[c]] for c in ("centroid_x", "centroid_y", "wsi_centroid_x", "wsi_centroid_y")
    }
    payload["wsi_polygon_hex"] = [None if p is None else [[float(x).hex(), float(y).hex()] for x, y in p]
                                  for p in out["wsi_polygon"]]
    payload["polygon_hex"] = [None if p is None else [[float(x).hex(), float(y).hex()] for x, y in p]
                              for p in nuc["polygon"]]
    payload["centroid_hex"] = [[float(a).hex(), float(b).hex()] for a, b in nuc["centroid"]]
    (OUT / "add_wsi_ref.json").write_text(json.dumps(payload))
    # error contract
    bad = nuc.copy()
    bad.loc[2, "tile_path"] = "/nowhere/patches/123_456.png"
    try:
        ref(bad, tiles)
        raise AssertionError("reference did not raise")
    except ValueError as e:
        (OUT / "add_wsi_ref_error.txt").write_text(str(e))


def make_graph():
    rng = np.random.default_rng(777)
    uni = rng.random((400, 2)) * 400.0
    gx, gy = np.meshgrid(np.arange(10.0) * 10.0 + 450.0, np.arange(10.0) * 10.0 + 30.0)
    lattice = np.stack([gx.ravel(), gy.ravel()], axis=1)
    dups = np.concatenate([uni[:60], lattice[:40]])
    coords = np.concatenate([uni, lattice, dups])
    types = rng.integers(1, 6, size=len(coords)).astype(np.int32)
    tree = cKDTree(coords)
    out = {"coords": coords, "types": types, "scipy_version": np.array(scipy.__version__)}
    for r in (10.0, 25.0, 40.0):
        pairs = tree.query_ball_tree(tree, r)  # ipynb:2964-2967
        edges = np.array([(i, j) for i, nb in enumerate(pairs) for j in nb if i != j and i < j], dtype=np.int64)
        edges = edges[np.lexsort((edges[:, 1], edges[:, 0]))]
        bf = graph.radius_graph_bruteforce(coords, r)
        assert np.array_equal(edges, bf["edges"])
        out[f"radius_{int(r)}_edges"] = edges
        out[f"radius_{int(r)}_dist"] = np.linalg.norm(coords[edges[:, 0]] - coords[edges[:, 1]], axis=1)  # ipynb:3041
    for k in (5, 8, 16):
        idx, dist = graph.knn(coords, k)
        bi, bd = graph.knn_bruteforce(coords, k)
        assert np.array_equal(idx, bi) and np.array_equal(dist, bd)
        # raw scipy output for the record (tie order is implementation-defined)
        sd, si = tree.query(coords, k + 1)
        out[f"knn_{k}_idx"] = idx
        out[f"knn_{k}_dist"] = dist
        out[f"knn_{k}_scipy_dist_sorted"] = np.sort(sd, axis=1)
        e, w, rp, col, ww = graph.undirected_union(idx, dist)
        g = graph.undirected_union_networkx(idx, dist)  # literal cell-11 loop
        assert g.number_of_edges() == len(e)
        for (a, b), wt in zip(e[:50], w[:50]):
            assert g.edges[int(a), int(b)]["weight"] == wt
        out[f"knn_{k}_und_edges"] = e
        out[f"knn_{k}_und_weight"] = w
    np.savez_compressed(OUT / "graph_small.npz", **out)


def make_known_answers():
    ka = {
        "source": "stored outputs of /root/reference/hovernet_tile_inference.ipynb (SURVEY Appendix D)",
        "centroids_yx": [[6.142857142857143, 211.26857142857142], [22.64962121212121, 296.1723484848485],
                         [28.736141906873613, 372.11529933481154], [30.28813559322034, 232.37853107344634],
                         [62.17614165890028, 206.6551724137931]],
        "distances": {"0,1": 86.493495, "0,3": 32.072182, "0,4": 56.222882, "1,2": 76.186465, "1,3": 64.249498,
                      "3,4": 40.969942},
        "counts": {"nodes": 101, "knn_k": 5, "knn_undirected_edges": 300, "knn_directed_edges": 505,
                   "filtered_nodes": 74, "filtered_edges": 174, "radius_um": 40.0, "mpp": 0.25, "radius_edges": 1189},
        "morph_columns": ["inst", "area", "perimeter", "major", "minor", "eccentricity", "perimeter_area",
                          "compactness", "roundness", "elongation"],
        "morph_rows": [
            [1, 175, 48.970563, 17.111364, 13.520166, 0.612942, 0.279832, 0.917018, 0.760990, 1.265618],
            [2, 528, 87.941125, 28.158967, 24.906426, 0.466552, 0.166555, 0.857946, 0.847834, 1.130590],
            [3, 451, 81.597980, 30.345906, 19.298996, 0.771716, 0.180927, 0.851192, 0.623572, 1.572409],
            [4, 177, 51.455844, 19.570430, 12.363558, 0.775175, 0.290711, 0.840067, 0.588414, 1.582912],
            [5, 1073, 124.225397, 44.577777, 30.756594, 0.723854, 0.115774, 0.873753, 0.687500, 1.449373],
            [97, 729, 145.189863, 49.893953, 21.443398, 0.902934, 0.199163, 0.434575, 0.372857, 2.326775],
            [98, 174, 58.384776, 25.002597, 9.440639, 0.925974, 0.335545, 0.641446, 0.354396, 2.648401],
            [99, 45, 25.313708, 10.387373, 6.082378, 0.810633, 0.562527, 0.882492, 0.531020, 1.707781],
            [100, 98, 44.727922, 19.587214, 6.899013, 0.935917, 0.456407, 0.615571, 0.325230, 2.839133],
            [101, 217, 56.627417, 22.495439, 12.395441, 0.834492, 0.260956, 0.850386, 0.545985, 1.814816],
        ],
    }
    (OUT / "notebook_known_answers.json").write_text(json.dumps(ka, indent=1))


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    make_add_wsi()
    make_graph()
    make_known_answers()
    print("fixtures written to", OUT)
