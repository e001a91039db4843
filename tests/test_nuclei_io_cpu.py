"""CPU: nuclei-table file I/O <-> SoA / CSR. The table formats are the ones the reference
writes with DataFrame.to_parquet / to_csv (aggregated_hovernet_run.py:398-402); the CUDA call is replaced by the
oracle-backed stand-in of test_host_cpu so the Arrow plumbing runs without a GPU."""
import numpy as np
import pandas as pd
import pytest

from conftest import frames_from_golden
from path_gene_multimodal_b200 import nuclei_io, synth
from test_host_cpu import _fake_map_morph_arrays


def _cells_equal(a, b):
    """Compare two list-valued cells as pandas hands them back from Parquet (ndarray of ndarrays) or as lists."""
    if a is None or b is None or (isinstance(a, float) and np.isnan(a)) or (isinstance(b, float) and np.isnan(b)):
        return (a is None or (isinstance(a, float) and np.isnan(a))) and (b is None or (isinstance(b, float) and np.isnan(b)))
    la = [list(map(float, v)) for v in a] if len(a) and np.ndim(a[0]) else list(map(float, a))
    lb = [list(map(float, v)) for v in b] if len(b) and np.ndim(b[0]) else list(map(float, b))
    return la == lb


def test_table_to_soa_reads_the_reference_parquet_layout(tmp_path, golden_add_wsi):
    nuc, tiles, expected = frames_from_golden(golden_add_wsi)
    path = tmp_path / "nuclei_local.parquet"
    nuc.to_parquet(path, index=False)                       # exactly what the reference does with its DataFrame
    table = nuclei_io.read_nuclei_table(path)
    soa = nuclei_io.table_to_soa(table)
    assert soa.n == len(nuc)
    assert np.array_equal(soa.centroid, np.array(nuc["centroid"].tolist()))
    assert np.array_equal(soa.bbox, np.array(nuc["bounding_box"].tolist()))
    assert soa.poly_is_none.tolist() == [p is None for p in nuc["polygon"]]
    want = [p for p in nuc["polygon"] if p is not None]
    assert soa.poly_off[-1] == sum(len(p) for p in want)
    assert np.array_equal(soa.poly_xy, np.array([v for p in want for v in p], dtype=np.float64).reshape(-1, 2))
    assert np.array_equal(np.diff(soa.poly_off), [0 if p is None else len(p) for p in nuc["polygon"]])
    assert np.array_equal(soa.type, nuc["type"].to_numpy())


def test_chunked_and_sliced_tables(tmp_path):
    import pyarrow as pa

    tab = synth.make_table(300, seed=5, dtype=np.float64)
    nuc, _ = synth.to_frames(tab)
    nuc.loc[[3, 77, 299], "polygon"] = None
    full = pa.Table.from_pandas(nuc, preserve_index=False)
    ref = nuclei_io.table_to_soa(full)
    chunked = pa.concat_tables([full.slice(0, 100), full.slice(100, 57), full.slice(157)])
    assert chunked["polygon"].num_chunks == 3
    got = nuclei_io.table_to_soa(chunked)
    for name in ("centroid", "bbox", "poly_off", "poly_xy", "poly_is_none", "type"):
        assert np.array_equal(getattr(got, name), getattr(ref, name)), name
    part = nuclei_io.table_to_soa(full.slice(50, 120))     # offsets of a slice do not start at 0
    assert part.poly_off[0] == 0 and np.array_equal(np.diff(part.poly_off), np.diff(ref.poly_off)[50:170])
    v0 = int(ref.poly_off[50])
    assert np.array_equal(part.poly_xy, ref.poly_xy[v0: v0 + int(part.poly_off[-1])])
    all_none = pa.table({"centroid": full["centroid"], "bounding_box": full["bounding_box"],
                         "polygon": pa.array([None] * 300)})
    soa = nuclei_io.table_to_soa(all_none)
    assert soa.poly_is_none.all() and soa.poly_off.tolist() == [0] * 301
    with pytest.raises(ValueError, match="centroid"):
        nuclei_io.fixed_list_column(pa.array([[1.0, 2.0], [3.0]]), 2, np.float64, "centroid")
    with pytest.raises(ValueError, match="pairs"):
        nuclei_io.polygon_column_to_csr(pa.array([[[1.0, 2.0, 3.0]]]))


def test_add_wsi_coords_to_table_matches_reference_golden(monkeypatch, tmp_path, golden_add_wsi):
    from path_gene_multimodal_b200 import nuclei_wsi

    monkeypatch.setattr(nuclei_wsi, "map_morph_arrays", _fake_map_morph_arrays)
    nuc, tiles, expected = frames_from_golden(golden_add_wsi)
    src = tmp_path / "local.parquet"
    nuc.to_parquet(src, index=False)
    out_pq, out_csv = tmp_path / "slide_hovernet_nuclei_wsi.parquet", tmp_path / "slide_hovernet_nuclei_wsi.csv"
    table = nuclei_io.process_nuclei_file(src, tiles, out_pq, out_csv)
    assert table.column_names == golden_add_wsi["out_columns"]
    # the reference's own files, written from its DataFrame, read back through pandas: ours must read back the same
    ref_pq, ref_csv = tmp_path / "ref.parquet", tmp_path / "ref.csv"
    expected.to_parquet(ref_pq, index=False)
    expected.to_csv(ref_csv, index=False)
    got, ref = pd.read_parquet(out_pq), pd.read_parquet(ref_pq)
    assert list(got.columns) == list(ref.columns)
    for c in ref.columns:
        if c in nuclei_io.LIST_COLUMNS:
            assert all(_cells_equal(a, b) for a, b in zip(got[c], ref[c])), c
        elif ref[c].dtype.kind in "fi":
            assert np.array_equal(got[c].to_numpy(), ref[c].to_numpy()) and got[c].dtype == ref[c].dtype, c
        else:
            assert got[c].tolist() == ref[c].tolist(), c
    assert out_csv.read_text() == ref_csv.read_text()       # the CSV twin, byte for byte
    # ... and the CSV twin parses back to the same arrays
    back = nuclei_io.table_to_soa(nuclei_io.read_nuclei_table(out_csv), polygon_col="wsi_polygon")
    direct = nuclei_io.table_to_soa(table, polygon_col="wsi_polygon")
    for name in ("centroid", "bbox", "poly_off", "poly_xy", "poly_is_none"):
        assert np.array_equal(getattr(back, name), getattr(direct, name)), name
    bad = nuc.copy()
    bad.loc[1, "tile_path"] = "/x/patches/9_9.png"
    import pyarrow as pa

    with pytest.raises(ValueError, match="Some nuclei have tile_key with no matching tile coords"):
        nuclei_io.add_wsi_coords_to_table(pa.Table.from_pandas(bad, preserve_index=False), tiles)


def test_table_path_equals_dataframe_path(monkeypatch):
    import pyarrow as pa
    from path_gene_multimodal_b200 import nuclei_wsi

    monkeypatch.setattr(nuclei_wsi, "map_morph_arrays", _fake_map_morph_arrays)
    tab = synth.make_table(500, seed=8, dtype=np.float64)
    nuc, tiles = synth.to_frames(tab)
    nuc.loc[[0, 10], "polygon"] = None
    df = nuclei_wsi.add_wsi_coords_to_nuclei(nuc, tiles, morphology=True)
    tb = nuclei_io.add_wsi_coords_to_table(pa.Table.from_pandas(nuc, preserve_index=False), tiles, morphology=True)
    assert tb.column_names == list(df.columns)
    got = tb.to_pandas()
    for c in df.columns:
        if c in nuclei_io.LIST_COLUMNS:
            assert all(_cells_equal(a, b) for a, b in zip(got[c], df[c])), c
        elif df[c].dtype.kind in "fi":
            assert np.array_equal(got[c].to_numpy(), df[c].to_numpy(), equal_nan=True), c
        else:
            assert got[c].tolist() == df[c].tolist(), c
    empty = nuclei_io.add_wsi_coords_to_table(pa.Table.from_pandas(nuc.iloc[:0], preserve_index=False), tiles)
    assert empty.num_rows == 0 and empty.column_names[-1] == "wsi_polygon"
