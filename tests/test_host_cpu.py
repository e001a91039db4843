"""CPU: host-side logic of the product (key join, list<->CSR conversion, column contract), the C-ABI
surface (library loads, exports every symbol the header declares) and the loud failure without a GPU.
The CUDA call is replaced by an oracle-backed stand-in here so the DataFrame plumbing can run on CPU."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT, frames_from_golden
from oracle import morphology as omorph
from oracle import tile_to_wsi as omap
from path_gene_multimodal_b200 import synth


def _fake_map_morph_arrays(poly_off, poly_xy, nuc_tile=None, tile_x=None, tile_y=None, centroid=None, bbox=None,
                           write_polygons=True, extra=False, device=None):
    wsi_c, wsi_b, wsi_p = omap.map_arrays(tile_x, tile_y, nuc_tile, centroid, bbox, poly_off, poly_xy)
    f = omorph.polygon_features_csr(poly_off, poly_xy)
    return {"wsi_centroid": wsi_c, "wsi_bbox": wsi_b.astype(np.int32), "wsi_poly_xy": wsi_p,
            "area": f["area"].astype(np.float32), "perimeter": f["perimeter"].astype(np.float32),
            "eccentricity": f["eccentricity"].astype(np.float32), "circularity": f["circularity"].astype(np.float32)}


def test_header_symbols_are_exported_and_bound():
    from path_gene_multimodal_b200 import _lib

    header = (ROOT / "include" / "pathgraph.h").read_text()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(pg_[a-z0-9_]+)\(", header, re.M))
    assert len(declared) >= 25
    lib = _lib.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in pathgraph.h but not exported by libpathgraph.so"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.pg_version() >= 100
    assert ctypes.sizeof(_lib.PgDegreeStats) == 32 and ctypes.sizeof(_lib.PgMorphOut) == 9 * 8


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from path_gene_multimodal_b200 import _lib, build_knn_graph, build_radius_graph, map_morph_arrays

    lib = _lib.load_library()
    h = ctypes.c_void_p()
    assert lib.pg_create(0, ctypes.byref(h)) != 0 and b"no CPU fallback" in lib.pg_last_error(None)
    for fn, args in ((build_knn_graph, ([[0.0, 0.0], [1.0, 1.0], [2.0, 0.0]], 1)), (build_radius_graph, ([[0.0, 0.0]], 1.0)),
                     (map_morph_arrays, (np.array([0, 3], dtype=np.int32), np.zeros((3, 2))))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn(*args)


def test_product_does_not_import_oracle():
    for p in (ROOT / "path_gene_multimodal_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{p} imports the oracle"
        if p.name == "interop.py":
            # the one export adapter: it may construct an nx.Graph object (the notebook's output type) from finished
            # arrays, but must not run any graph algorithm or query of networkx
            assert not re.search(r"\bnx\.(?!Graph\b)\w+", src), f"{p} uses networkx for more than constructing nx.Graph"
            assert not re.search(r"^\s*(from|import)\s+(scipy|sklearn)\b", src, re.M)
            continue
        assert not re.search(r"^\s*(from|import)\s+(scipy|networkx|sklearn)\b", src, re.M), f"{p} imports a CPU graph library"


def test_polygons_csr_roundtrip():
    from path_gene_multimodal_b200.nuclei_wsi import csr_to_polygons, polygons_to_csr

    polys = pd.Series([[[0.5, 1.0], [2.0, 3.5], [4.0, 0.0]], None, [], [[1.0, 1.0], [2.0, 2.0], [3.0, 1.0], [1.0, 1.0]]], dtype=object)
    off, xy, is_none = polygons_to_csr(polys)
    assert off.tolist() == [0, 3, 3, 3, 7] and xy.shape == (7, 2) and is_none.tolist() == [False, True, False, False]
    back = csr_to_polygons(off, xy, is_none)
    assert back[0] == polys[0] and back[1] is None and back[2] == [] and back[3] == polys[3]
    with pytest.raises(Exception):
        polygons_to_csr(pd.Series([[[1.0, 2.0, 3.0]]], dtype=object))


def test_add_wsi_coords_host_logic_against_reference_golden(monkeypatch, golden_add_wsi):
    from path_gene_multimodal_b200 import nuclei_wsi

    monkeypatch.setattr(nuclei_wsi, "map_morph_arrays", _fake_map_morph_arrays)
    nuc, tiles, expected = frames_from_golden(golden_add_wsi)
    before = nuc.copy(deep=True)
    got = nuclei_wsi.add_wsi_coords_to_nuclei(nuc, tiles)
    assert list(got.columns) == golden_add_wsi["out_columns"]
    for c in golden_add_wsi["out_columns"]:
        if got[c].dtype.kind in "fi":
            assert np.array_equal(got[c].to_numpy(), expected[c].to_numpy()), c
            assert str(got[c].dtype) == golden_add_wsi["out_dtypes"][c], c
        else:
            assert got[c].tolist() == expected[c].tolist(), c
    pd.testing.assert_frame_equal(nuc, before)
    nuc.loc[2, "tile_path"] = "/nowhere/patches/123_456.png"
    with pytest.raises(ValueError, match="Some nuclei have tile_key with no matching tile coords"):
        nuclei_wsi.add_wsi_coords_to_nuclei(nuc, tiles)


def test_add_wsi_coords_host_logic_morphology_and_custom_keys(monkeypatch):
    from path_gene_multimodal_b200 import nuclei_wsi

    monkeypatch.setattr(nuclei_wsi, "map_morph_arrays", _fake_map_morph_arrays)
    tab = synth.make_table(120, seed=2, dtype=np.float64)
    nuc, tiles = synth.to_frames(tab)
    nuc = nuc.rename(columns={"tile_path": "src"})
    tiles = tiles.rename(columns={"png_path": "file"})
    got = nuclei_wsi.add_wsi_coords_to_nuclei(nuc, tiles, tile_key_col_nuc="src", tile_key_col_tiles="file", morphology=True)
    exp = omap.add_wsi_coords_to_nuclei_oracle(nuc, tiles, "src", "file")
    pd.testing.assert_frame_equal(got[exp.columns], exp)
    assert {"area", "perimeter", "eccentricity", "circularity"} <= set(got.columns)


def test_synth_generator_shapes():
    tab = synth.make_table(1000, seed=synth.SEEDS["C1"])
    assert tab.poly_xy.dtype == np.float32 and np.all(tab.poly_xy * 2 == np.round(tab.poly_xy * 2))   # 0.5-px lattice
    nv = np.diff(tab.poly_off)
    assert nv.min() >= 8 and nv.max() <= 32 and set(np.unique(tab.types)) <= {1, 2, 3, 4, 5}
    xy, types, side = synth.make_points(100_000, synth.SEEDS["C1"])
    assert abs(len(xy) / side ** 2 - synth.DENSITY) / synth.DENSITY < 0.05 and side % 508 == 0
    off, pxy = synth.make_polygons(10, 1, v_fixed=32)
    assert np.all(np.diff(off) == 32)


def test_bench_reference_pipeline_small():
    import bench

    xy, types, _ = synth.make_points(3000, 5)
    ne, ei, ea = bench.reference_pipeline(xy, types, 50.0)
    from oracle import graph as ograph

    assert ne == len(ograph.radius_graph(xy, 50.0)["edges"]) and ei == (4, ne) and ea == (2 * ne, 1)
    assert bench.algorithmic_bytes(10, 4) == 512


def test_halfpx_staging_host_side():
    """cohort.to_halfpx: exact on the half-pixel lattice find_contours emits (aggregated_hovernet_run.py:185), refusing
    anything it would have to round; the synthetic tables are on that lattice, so the staging halves their polygon bytes."""
    from path_gene_multimodal_b200 import cohort, synth

    off, xy = synth.make_polygons(500, 5)
    q = cohort.to_halfpx(xy)
    assert q.dtype == np.int16 and q.shape == xy.shape and q.nbytes * 2 == xy.nbytes
    assert np.array_equal(q.astype(np.float32) * np.float32(0.5), xy)          # what pg_widen_halfpx computes
    assert np.array_equal(cohort.to_halfpx(xy.astype(np.float64)), q)
    for bad in ([[0.25, 1.0]], [[1.0, np.nan]], [[16384.0, 0.0]], [[-16384.0, 0.0]]):
        with pytest.raises(ValueError):
            cohort.to_halfpx(np.array(bad))
    assert np.array_equal(cohort.to_halfpx(np.array([[16383.5, -16383.5]])), np.array([[32767, -32767]], dtype=np.int16))
    with pytest.raises(ValueError):
        cohort.pin_table(None, staging="int8")
