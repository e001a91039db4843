"""GPU tests of the round-2 additions: compact radius outputs, the rest of the cell-11 surface (neighbour
coordinates, nx.Graph keyed by nuc_id), centroid_order="yx", numpy-array table cells, the per-axis bbox dtype,
float64 ring length on long rings, C3 parity at 250 k polygons, two engines in one process."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import graph as ograph
from oracle import morphology as omorph
from oracle import tile_to_wsi as omap
from path_gene_multimodal_b200 import synth

pytestmark = pytest.mark.gpu


def test_radius_compact_outputs_equal_notebook_outputs(golden_graph):
    from path_gene_multimodal_b200 import build_radius_graph, edge_index_from_edges

    coords, types = golden_graph["coords"], golden_graph["types"]
    full = build_radius_graph(coords, r=25.0, types=types)
    for _ in range(3):  # first call: exact count -> fill; later calls: the one-enqueue path with a capacity hint
        c = build_radius_graph(coords, r=25.0, types=types, outputs="compact")
        assert c["edges"].dtype == np.int32 and np.array_equal(c["edges"], full["edges"])
        assert c["dist"].dtype == np.float32 and np.array_equal(c["dist"], full["dist"])
        assert np.array_equal(c["degree"], full["degree"]) and np.array_equal(c["nbr_count"], full["nbr_count"])
        assert c["degree_stats"]["sum"] == full["degree_stats"]["sum"]
        ei, ea = edge_index_from_edges(c["edges"], c["dist"])
        assert ei.dtype == np.int64 and np.array_equal(ei, full["edge_index"]) and np.array_equal(ea, full["edge_attr"])
    # a denser slide after a sparse one: the hint is too small once, the call must still be exact
    xy, ty, _ = synth.make_points(30_000, seed=9)
    a = build_radius_graph(xy, r=20.0, types=ty, outputs="compact")
    b = build_radius_graph(xy, r=120.0, types=ty, outputs="compact")
    ref = ograph.radius_graph(xy, 120.0)
    assert np.array_equal(b["edges"], ref["edges"]) and len(a["edges"]) < len(b["edges"])
    t_ei, t_ea = edge_index_from_edges(torch.from_numpy(b["edges"]).cuda(), torch.from_numpy(b["dist"]).cuda())
    assert t_ei.dtype == torch.int64 and np.array_equal(t_ei.cpu().numpy(), ref["edge_index"])
    assert np.array_equal(t_ea.cpu().numpy(), ref["edge_attr"])


def test_knn_graph_frame_is_cell_11(golden_graph):
    import networkx as nx

    from path_gene_multimodal_b200 import knn_graph_frame, to_networkx

    coords, types = golden_graph["coords"][:400], golden_graph["types"][:400]
    n = len(coords)
    final_df = pd.DataFrame({
        "nuc_id": [f"{i:08x}" for i in range(n)],
        "centroid": [[float(y), float(x)] for x, y in coords],                 # HoverNeXt order: [y, x]
        "type_name": [{1: "neoplastic", 2: "inflammatory", 3: "connective", 4: "dead", 5: "epithelial"}[int(t)] for t in types],
    })
    df, g = knn_graph_frame(final_df, k=5)
    idx, dist = ograph.knn(np.ascontiguousarray(coords), 5)
    assert df["knn_neighbors"].tolist() == idx.tolist()
    assert df["knn_neighbor_distances"].tolist() == dist.tolist()
    assert df["knn_neighbor_coords"].iloc[7] == [(float(coords[j, 0]), float(coords[j, 1])) for j in idx[7]]
    assert df["type_id"].tolist() == [int(t) for t in types] and df["color"].iloc[0].startswith("tab:")
    assert "centroid" in final_df.columns and "knn_neighbors" not in final_df.columns  # the input is not modified
    G = to_networkx(g, df)
    # the literal cell-11 loop (networkx has_edge / add_edge), keyed by nuc_id
    ref = nx.Graph()
    ids = final_df["nuc_id"].tolist()
    ref.add_nodes_from(ids)
    for i in range(n):
        for j, d in zip(idx[i], dist[i]):
            a, b = ids[i], ids[int(j)]
            if ref.has_edge(a, b):
                ref.edges[a, b]["weight"] = min(ref.edges[a, b]["weight"], float(d))
            else:
                ref.add_edge(a, b, weight=float(d))
    assert set(G.nodes) == set(ref.nodes) and G.number_of_edges() == ref.number_of_edges()
    assert all(G.edges[u, v]["weight"] == ref.edges[u, v]["weight"] for u, v in ref.edges)
    assert G.nodes[ids[3]]["type_id"] == int(types[3]) and G.nodes[ids[3]]["pos"] == (float(coords[3, 0]), float(coords[3, 1]))


def test_centroid_order_yx_and_array_cells():
    from path_gene_multimodal_b200 import add_wsi_coords_to_nuclei

    tab = synth.make_table(500, seed=31, dtype=np.float64)
    nuc, tiles = synth.to_frames(tab)
    ref = add_wsi_coords_to_nuclei(nuc, tiles)
    fix = add_wsi_coords_to_nuclei(nuc, tiles, centroid_order="yx")
    cen = np.array(nuc["centroid"].tolist())
    assert np.array_equal(fix["centroid_x"], cen[:, 1]) and np.array_equal(fix["centroid_y"], cen[:, 0])
    assert np.array_equal(fix["wsi_centroid_x"], fix["tile_x"] + cen[:, 1])
    assert np.array_equal(fix["wsi_centroid_y"], fix["tile_y"] + cen[:, 0])
    assert fix["wsi_polygon"].tolist() == ref["wsi_polygon"].tolist() and np.array_equal(fix["wsi_bbox_xmin"], ref["wsi_bbox_xmin"])
    with pytest.raises(ValueError):
        add_wsi_coords_to_nuclei(nuc, tiles, centroid_order="zz")
    # cells as numpy arrays (what pd.read_parquet hands back for list columns)
    arr = nuc.copy()
    arr["polygon"] = [None if p is None else np.array([np.array(v) for v in p], dtype=object) for p in nuc["polygon"]]
    arr["centroid"] = [np.array(c) for c in nuc["centroid"]]
    arr["bounding_box"] = [np.array(b) for b in nuc["bounding_box"]]
    arr.loc[arr.index[5], "polygon"] = None
    nuc2 = nuc.copy()
    nuc2.loc[nuc2.index[5], "polygon"] = None
    got, want = add_wsi_coords_to_nuclei(arr, tiles), add_wsi_coords_to_nuclei(nuc2, tiles)
    assert got["wsi_polygon"].tolist() == want["wsi_polygon"].tolist() and got["wsi_polygon"].iloc[5] is None
    assert np.array_equal(got["wsi_centroid_x"], want["wsi_centroid_x"])


def test_wsi_bbox_dtype_follows_each_axis():
    from path_gene_multimodal_b200 import add_wsi_coords_to_nuclei

    tab = synth.make_table(64, seed=32, dtype=np.float64)
    nuc, tiles = synth.to_frames(tab)
    tiles = tiles.copy()
    tiles["x"] = tiles["x"].astype(np.float64)          # e.g. after a CSV round trip; y stays int64
    out = add_wsi_coords_to_nuclei(nuc, tiles)
    exp = omap.add_wsi_coords_to_nuclei_oracle(nuc, tiles)
    for c in ("wsi_bbox_xmin", "wsi_bbox_ymin", "wsi_bbox_xmax", "wsi_bbox_ymax", "tile_x", "tile_y"):
        assert out[c].dtype == exp[c].dtype, c
        assert np.array_equal(out[c].to_numpy(), exp[c].to_numpy()), c


def test_ring_length_is_float64_accurate_on_long_rings():
    from path_gene_multimodal_b200 import polygon_morphology_table

    rng = np.random.default_rng(8)
    rings = []
    for nv in (50_000, 6_000, 900, 33, 5):
        t = np.sort(rng.uniform(0, 2 * np.pi, nv))
        rad = 4000.0 * (1.0 + 0.3 * np.sin(7 * t)) + rng.uniform(-3, 3, nv)
        rings.append(np.round(np.stack([20000 + rad * np.cos(t), 15000 + rad * np.sin(t)], axis=1) * 2) / 2)
    off = np.concatenate([[0], np.cumsum([len(r) for r in rings])]).astype(np.int32)
    xy = np.concatenate(rings)
    tab = polygon_morphology_table(xy, poly_off=off)
    ref = omorph.polygon_features_csr(off, xy)
    np.testing.assert_allclose(tab["perimeter_px"], ref["perimeter"], rtol=1e-6)   # float32 output of a float64 sum
    np.testing.assert_allclose(tab["area_px2"], ref["area"], rtol=1e-6)
    np.testing.assert_allclose(tab["centroid_x"], ref["centroid_x"], rtol=1e-9)


def test_c3_parity_at_250k_polygons(engine):
    n = 250_000
    off, xy = synth.make_polygons(n, synth.SEEDS["C3"], v_fixed=32)
    res = engine.map_morph(torch.from_numpy(off).cuda(), torch.from_numpy(xy).cuda(), write_polygons=False, extra=True)
    ref = omorph.polygon_features_csr(off, xy)
    for name, key in (("area", "area"), ("perimeter", "perimeter"), ("circularity", "circularity"),
                      ("major_axis", "major_axis_length")):
        np.testing.assert_allclose(res[name].cpu().numpy(), ref[key], rtol=1e-5, err_msg=name)
    np.testing.assert_allclose(res["minor_axis"].cpu().numpy(), ref["minor_axis_length"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(res["eccentricity"].cpu().numpy(), ref["eccentricity"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(res["centroid_x"].cpu().numpy(), ref["centroid_x"], rtol=1e-9)


def test_two_engines_in_one_process_share_nothing():
    # K1's shared-memory opt-in: cudaFuncSetAttribute acts on one device and SETS the limit. A process-wide static
    # (round 1) skipped the opt-in on a second device; a per-handle "largest size so far" let a second handle on the same
    # device LOWER the limit under the first one. Every handle now opts in once for the one maximum: big slab on A,
    # small slab on B (same device, or another one when there is one), big slab on A again.
    from path_gene_multimodal_b200.engine import Engine

    big = synth.make_table(2000, seed=33, dtype=np.float64, v_lo=60, v_hi=90)      # long rings: a slab well above 48 KB
    small = synth.make_table(3000, seed=34, dtype=np.float64, v_lo=8, v_hi=12)
    ref_big = omorph.polygon_features_csr(big.poly_off, big.poly_xy)
    ref_small = omorph.polygon_features_csr(small.poly_off, small.poly_xy)
    dev_b = 1 if torch.cuda.device_count() > 1 else 0
    eng_a, eng_b = Engine(0), Engine(dev_b)

    def run(eng, d, tab, ref):
        with torch.cuda.device(d):
            res = eng.map_morph(torch.from_numpy(tab.poly_off).cuda(d), torch.from_numpy(tab.poly_xy).cuda(d), write_polygons=False)
            np.testing.assert_allclose(res["area"].cpu().numpy(), ref["area"], rtol=1e-5)
            np.testing.assert_allclose(res["perimeter"].cpu().numpy(), ref["perimeter"], rtol=1e-5)

    run(eng_a, 0, big, ref_big)
    run(eng_b, dev_b, small, ref_small)
    run(eng_a, 0, big, ref_big)
    run(eng_b, dev_b, big, ref_big)
    eng_a.close()
    eng_b.close()


def test_cohort_runner_digest_does_not_depend_on_lanes():
    from path_gene_multimodal_b200 import cohort

    n_slides, n = 5, 20_000
    cache, tables = {}, {}
    for s in range(n_slides):
        tables[s] = cohort.pin_table(synth.make_cohort_slide(s, n, n_bases=2), cache)
    assert tables[0].poly_xy is tables[2].poly_xy and tables[0].nuc_tile is not tables[2].nuc_tile   # bases shared, deals not
    digests = []
    for lanes in (1, 2, 3):
        runner = cohort.CohortRunner(0, lanes=lanes)
        res, ms = runner.run(range(n_slides), tables.__getitem__)
        runner.close()
        assert sorted(res) == list(range(n_slides)) and ms > 0
        digests.append(cohort.cohort_checksum(res))
    assert digests[0] == digests[1] == digests[2]
    # one slide against the oracle: radius edges and kNN union edges of its WSI centroids
    coords = tables[3].wsi_centroids()
    ref = ograph.radius_graph(coords, 50.0)
    idx, dist = ograph.knn(coords, 8)
    e, w, rp, col, _ = ograph.undirected_union(idx, dist)
    one = cohort.process_slide(cohort.get_engine(0), tables[3])
    assert one["radius_edges"] == len(ref["edges"]) and one["knn_edges"] == len(e)
    assert one["radius_edge_hash"] == int((ref["edges"][:, 0] * 1000003 + ref["edges"][:, 1]).sum())
    assert one["knn_edge_hash"] == int((e[:, 0] * 1000003 + e[:, 1]).sum())


def test_knn_union_presized_needs_no_totals(engine):
    from path_gene_multimodal_b200.engine import default_knn_cell

    xy, types, side = synth.make_points(40_000, 91)
    d_xy, d_ty = torch.from_numpy(xy).cuda(), torch.from_numpy(types).cuda()
    engine.grid_build(d_xy, d_ty, None, default_knn_cell(len(xy), float(side) ** 2, 8), None)
    kn = engine.knn(8, dist_dtype=torch.float32)
    a = engine.knn_union(kn["knn_idx"], kn["dist32"], types=d_ty, n_types=5, symmetric_dist=True)
    b = engine.knn_union(kn["knn_idx"], kn["dist32"], types=d_ty, n_types=5, symmetric_dist=True, presized=True)
    e, eu = int(a["row_ptr"][-1]), int(a["up_ptr"][-1])
    assert b["col"].shape[0] == 2 * 8 * len(xy) and int(b["row_ptr"][-1]) == e and int(b["up_ptr"][-1]) == eu
    assert torch.equal(b["row_ptr"], a["row_ptr"]) and torch.equal(b["col"][:e], a["col"]) and torch.equal(b["w"][:e], a["w"])
    assert torch.equal(b["edges"][:eu], a["edges"]) and torch.equal(b["edge_w"][:eu], a["edge_w"])
    assert torch.equal(b["degree"], a["degree"]) and torch.equal(b["nbr_count"], a["nbr_count"])


def test_table_pass_replays_as_one_cuda_graph(engine):
    from path_gene_multimodal_b200.engine import default_knn_cell

    tab = synth.make_table(30_000, seed=41)
    side_px = float(tab.n_tiles_side * 508)
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(getattr(tab, k)).to(dev) for k in ("poly_off", "poly_xy", "nuc_tile", "tile_x", "tile_y", "centroid", "bbox", "types")}
    keep = {}

    def table_pass():
        mm = engine.map_morph(t["poly_off"], t["poly_xy"], t["nuc_tile"], t["tile_x"], t["tile_y"], t["centroid"], t["bbox"], out=keep.get("mm"))
        keep["mm"] = mm
        engine.grid_build(mm["wsi_centroid"], t["types"], None, default_knn_cell(tab.n, side_px ** 2, 8), (0.0, 0.0, side_px, side_px))
        keep["kn"] = engine.knn(8, dist_dtype=torch.float32, out=keep.get("kn"))
        keep["un"] = engine.knn_union(keep["kn"]["knn_idx"], keep["kn"]["dist32"], types=t["types"], symmetric_dist=True, presized=True)

    table_pass()
    torch.cuda.synchronize()
    want = {k: keep["un"][k].clone() for k in ("row_ptr", "col", "edges", "degree", "nbr_count")}
    graph = engine.capture(table_pass)
    for k in ("col", "edges", "degree"):
        keep["un"][k].zero_()
    graph.replay()
    torch.cuda.synchronize()
    e, eu = int(want["row_ptr"][-1]), int(keep["un"]["up_ptr"][-1])
    assert torch.equal(keep["un"]["row_ptr"], want["row_ptr"]) and torch.equal(keep["un"]["col"][:e], want["col"][:e])
    assert torch.equal(keep["un"]["edges"][:eu], want["edges"][:eu]) and torch.equal(keep["un"]["degree"], want["degree"])
    idx, dist = ograph.knn(tab.wsi_centroids(), 8)
    ref_e = ograph.undirected_union(idx, dist)[0]
    assert np.array_equal(keep["un"]["edges"][:eu].cpu().numpy(), ref_e)


def test_compact_narrow_counts(golden_graph):
    from path_gene_multimodal_b200 import build_radius_graph

    coords, types = golden_graph["coords"], golden_graph["types"]
    full = build_radius_graph(coords, r=25.0, types=types)
    for dt in (np.uint8, np.uint16):
        for _ in range(2):
            c = build_radius_graph(coords, r=25.0, types=types, outputs="compact", count_dtype=dt)
            assert c["degree"].dtype == dt and c["nbr_count"].dtype == dt
            assert np.array_equal(c["degree"], full["degree"]) and np.array_equal(c["nbr_count"], full["nbr_count"])
            assert np.array_equal(c["edges"], full["edges"])
    # 300 coincident points: degree 299 does not fit uint8 -> OverflowError, and the next call is clean again
    dense = np.concatenate([np.zeros((300, 2)), coords[:100] + 1e4])
    with pytest.raises(OverflowError):
        build_radius_graph(dense, r=1.0, types=np.ones(len(dense), dtype=np.int32), outputs="compact", count_dtype=np.uint8)
    ok = build_radius_graph(dense, r=1.0, types=np.ones(len(dense), dtype=np.int32), outputs="compact", count_dtype=np.uint16)
    assert int(ok["degree"].max()) == 299 and int(ok["nbr_count"][0, 0]) == 299
    with pytest.raises(ValueError):
        build_radius_graph(coords, r=25.0, count_dtype=np.uint8)


def test_c4_slide_full_size_against_scipy():
    """One slide of BASELINE config 4 at full size (500 k nuclei, ragged rings) through the cohort's slide pass; the
    graphs against scipy (radius edges, kNN-8 lists -> union edges), bit-exact."""
    from path_gene_multimodal_b200 import cohort

    tab = cohort.pin_table(synth.make_cohort_slide(7, 500_000))
    one = cohort.process_slide(cohort.get_engine(0), tab)
    coords = tab.wsi_centroids()
    ref = ograph.radius_graph(coords, 50.0)
    assert one["radius_edges"] == len(ref["edges"])
    assert one["radius_edge_hash"] == int((ref["edges"][:, 0] * 1000003 + ref["edges"][:, 1]).sum())
    assert one["nbr_sum"] == int(2 * len(ref["edges"]))          # every type is in 1..5: the type counts sum to the degrees
    idx, dist = ograph.knn(coords, 8)
    e = ograph.undirected_union(idx, dist)[0]
    assert one["knn_edges"] == len(e) and one["knn_edge_hash"] == int((e[:, 0] * 1000003 + e[:, 1]).sum())
    assert one["knn_deg_sum"] == 2 * len(e)
    feat = omorph.polygon_features_csr(tab.poly_off, tab.poly_xy)
    np.testing.assert_allclose(one["area_sum"], float(np.float32(feat["area"]).astype(np.float64).sum()), rtol=1e-6)


def test_wsi_polygon_as_arrow_is_the_same_column():
    from path_gene_multimodal_b200 import add_wsi_coords_to_nuclei

    tab = synth.make_table(400, seed=35, dtype=np.float64)
    nuc, tiles = synth.to_frames(tab)
    nuc.loc[nuc.index[3], "polygon"] = None
    ref = add_wsi_coords_to_nuclei(nuc, tiles)
    arw = add_wsi_coords_to_nuclei(nuc, tiles, wsi_polygon_as="arrow")
    assert isinstance(arw["wsi_polygon"].dtype, pd.ArrowDtype)
    got = [None if v is None or v is pd.NA else [list(p) for p in v] for v in arw["wsi_polygon"].tolist()]
    assert got == ref["wsi_polygon"].tolist() and got[3] is None
    # and it feeds straight back in (its buffers are the CSR): the morphology of the shifted rings
    again = nuc.drop(columns=["polygon"]).assign(polygon=arw["wsi_polygon"])
    out = add_wsi_coords_to_nuclei(again, tiles, morphology=True)
    base = add_wsi_coords_to_nuclei(nuc, tiles, morphology=True)
    np.testing.assert_allclose(out["area"].to_numpy(), base["area"].to_numpy(), rtol=1e-6, equal_nan=True)


def test_round2_entry_points_on_empty_and_tiny_inputs(engine):
    from path_gene_multimodal_b200 import build_radius_graph, knn_graph_frame

    dev = torch.device("cuda", 0)
    e_xy = torch.zeros((0, 2), dtype=torch.float64, device=dev)
    e_i = torch.zeros((0,), dtype=torch.int32, device=dev)
    recs, totals = engine.strip_partition(e_xy, e_i, e_i, [10.0, 20.0])
    assert recs.shape[0] == 0 and totals.tolist() == [0, 0, 0]
    xy_all = torch.zeros((4, 2), dtype=torch.float64, device=dev)
    counts = engine.halo_unpack_multi(torch.zeros((0, 3), dtype=torch.float64, device=dev), 0, 0, 0, [(0.0, 1.0), (1.0, 2.0)],
                                      xy_all, torch.zeros(4, dtype=torch.int32, device=dev), torch.zeros(4, dtype=torch.int32, device=dev), 0)
    assert counts.tolist() == [0, 0]
    id_map, tbg = engine.gid_maps(e_i, e_i, 0, 5)
    assert id_map.tolist() == [-1] * 5 and tbg.tolist() == [0] * 5
    assert engine.narrow_counts(torch.zeros((0, 5), dtype=torch.int32, device=dev), torch.uint8).shape == (0, 5)
    odd = torch.arange(7, dtype=torch.int32, device=dev) * 40            # not a multiple of four: the tail path; 240 fits, 280 would not
    assert engine.narrow_counts(odd, torch.uint8).tolist() == [0, 40, 80, 120, 160, 200, 240]
    engine.check_overflow()
    assert engine.knn_neighbor_coords(torch.zeros((0, 3), dtype=torch.int32, device=dev), torch.zeros((5, 2), dtype=torch.float64, device=dev)).shape == (0, 3, 2)
    # public API on nothing / one point / two points, both output sets
    for outputs in ("notebook", "compact"):
        g0 = build_radius_graph(np.zeros((0, 2)), r=5.0, types=np.zeros(0, dtype=np.int32), outputs=outputs)
        assert g0["edges"].shape == (0, 2) and g0["degree"].shape == (0,)
        g1 = build_radius_graph(np.array([[1.0, 2.0]]), r=5.0, types=np.array([3]), outputs=outputs)
        assert g1["edges"].shape == (0, 2) and g1["degree"].tolist() == [0] and g1["nbr_count"].tolist() == [[0, 0, 0, 0, 0]]
        g2 = build_radius_graph(np.array([[0.0, 0.0], [3.0, 4.0]]), r=5.0, types=np.array([1, 2]), outputs=outputs)
        assert g2["edges"].tolist() == [[0, 1]] and g2["dist"].tolist() == [5.0] and g2["nbr_count"].tolist() == [[0, 1, 0, 0, 0], [1, 0, 0, 0, 0]]
    # cell 11 with a `type` column instead of `type_name`, k = N - 1
    df = pd.DataFrame({"nuc_id": list("abcd"), "centroid": [[0.0, 0.0], [0.0, 1.0], [5.0, 0.0], [5.0, 1.5]], "type": [1, 2, 7, 5]})
    out, g = knn_graph_frame(df, k=3)
    assert out["knn_neighbors"].tolist()[0] == [1, 2, 3] and out["color"].tolist() == ["tab:red", "tab:green", "black", "tab:orange"]
    assert out["knn_neighbor_coords"].iloc[0][0] == (1.0, 0.0)          # centroid cells are [y, x]
    with pytest.raises(ValueError):
        knn_graph_frame(df, k=4)


def test_halfpx_staging_is_exact_and_changes_nothing(engine):
    """Contour vertices staged as int16 half-pixels (cohort.to_halfpx) widen to the very same float32 values
    (pg_widen_halfpx), so a cohort slide gives the same summary either way; off-lattice vertices are refused."""
    from path_gene_multimodal_b200 import cohort, synth
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(11)
    for m in (0, 1, 7, 8, 9, 4099):
        q = rng.integers(-32768, 32768, size=(m, 2), dtype=np.int64).astype(np.int16)
        got = engine.widen_halfpx(torch.from_numpy(q).to(dev)).cpu().numpy()
        assert got.dtype == np.float32 and np.array_equal(got, q.astype(np.float32) * np.float32(0.5))
    with pytest.raises(ValueError):
        cohort.to_halfpx(np.array([[0.25, 1.0]]))
    with pytest.raises(ValueError):
        cohort.to_halfpx(np.array([[20000.0, 1.0]]))
    n = 60_000
    a = cohort.pin_table(synth.make_cohort_slide(3, n), {})
    b = cohort.pin_table(synth.make_cohort_slide(3, n), {}, staging="halfpx16")
    assert a.poly_xy.dtype == np.float32 and b.poly_xy.dtype == np.int16 and b.poly_xy.nbytes * 2 == a.poly_xy.nbytes
    ra, rb = cohort.process_slide(engine, a), cohort.process_slide(engine, b)
    assert ra == rb and ra["knn_edges"] > 0 and ra["radius_edges"] > 0
