"""Pins on an independent third-party engine: OpenCV (cv2 4.13 is in this image, on the GPU box as well).

shapely / skimage - the libraries the reference calls for a4 / a5 / f3 - are absent here AND on the GPU box
(profiles/r2/probe_libs_gpu_box.json), so the oracle for those rows is a restatement of their published formulas.
OpenCV implements the same mathematics independently (Green's-theorem polygon area / moments: cv2.contourArea,
cv2.moments on a point contour; ring length: cv2.arcLength; raster moments: cv2.moments on a binary image), so it
pins both the restated oracle (CPU tests below) and the CUDA kernels (gpu tests below) on numbers neither of them
produced. What stays unpinned: skimage's perimeter weighting, convex_hull_image and find_contours /
approximate_polygon tie handling (OpenCV's counterparts are different algorithms).
"""
import numpy as np
import pytest

try:
    import cv2
except ImportError:  # the committed OpenCV outputs (tests/golden/opencv_pins.npz) still pin everything below
    cv2 = None
needs_cv2 = pytest.mark.skipif(cv2 is None, reason="OpenCV not installed (the committed-vector tests still run)")

from pathlib import Path  # noqa: E402

from oracle import morphology as omorph  # noqa: E402
from oracle import raster as oraster  # noqa: E402
from path_gene_multimodal_b200 import synth  # noqa: E402


def cv2_polygon_features(off, xy):
    """area, perimeter, centroid and second-moment eccentricity / axes of every ring, from OpenCV alone."""
    n = len(off) - 1
    out = {k: np.full(n, np.nan) for k in ("area", "perimeter", "centroid_x", "centroid_y", "eccentricity",
                                           "major_axis_length", "minor_axis_length")}
    for i in range(n):
        p = np.ascontiguousarray(xy[off[i]:off[i + 1]], dtype=np.float32)
        if len(p) < 3:
            continue
        # frame at the first vertex: keeps cv2's float32 points exact for half-pixel lattices far from the origin
        q = (p.astype(np.float64) - p[0].astype(np.float64)).astype(np.float32)
        out["area"][i] = cv2.contourArea(q)
        out["perimeter"][i] = cv2.arcLength(q, True)
        m = cv2.moments(q)
        if m["m00"] == 0:
            continue
        out["centroid_x"][i] = m["m10"] / m["m00"] + float(p[0, 0])
        out["centroid_y"][i] = m["m01"] / m["m00"] + float(p[0, 1])
        mu20, mu02, mu11 = m["mu20"] / m["m00"], m["mu02"] / m["m00"], m["mu11"] / m["m00"]
        mm, cc = 0.5 * (mu20 + mu02), np.hypot(0.5 * (mu20 - mu02), mu11)
        l1, l2 = mm + cc, max(mm - cc, 0.0)
        out["eccentricity"][i] = np.sqrt(1.0 - l2 / l1) if l1 > 0 else 0.0
        out["major_axis_length"][i] = 4.0 * np.sqrt(max(l1, 0.0))
        out["minor_axis_length"][i] = 4.0 * np.sqrt(l2)
    return out


def cv2_raster_props(inst_map, n_labels):
    out = {k: np.full(n_labels, np.nan) for k in ("area", "centroid_r", "centroid_c", "eccentricity", "major_axis_length",
                                                  "minor_axis_length")}
    bbox = np.zeros((n_labels, 4), dtype=np.int64)
    for l in range(1, n_labels + 1):
        mask = (inst_map == l).astype(np.uint8)
        if not mask.any():
            continue
        m = cv2.moments(mask, binaryImage=True)          # x = column, y = row
        x, y, w, h = cv2.boundingRect(mask)
        bbox[l - 1] = (y, x, y + h, x + w)
        out["area"][l - 1] = m["m00"]
        out["centroid_r"][l - 1] = m["m01"] / m["m00"]
        out["centroid_c"][l - 1] = m["m10"] / m["m00"]
        mu_rr, mu_cc, mu_rc = m["mu02"] / m["m00"], m["mu20"] / m["m00"], m["mu11"] / m["m00"]
        mm, cc = 0.5 * (mu_rr + mu_cc), np.hypot(0.5 * (mu_rr - mu_cc), mu_rc)
        l1, l2 = mm + cc, max(mm - cc, 0.0)
        out["eccentricity"][l - 1] = np.sqrt(1.0 - l2 / l1) if l1 > 0 else 0.0
        out["major_axis_length"][l - 1] = 4.0 * np.sqrt(l1)
        out["minor_axis_length"][l - 1] = 4.0 * np.sqrt(l2)
    out["bbox"] = bbox
    return out


def blobs(h, w, n, seed):
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), dtype=np.int32)
    yy, xx = np.mgrid[0:h, 0:w]
    for l in range(1, n + 1):
        cy, cx = rng.uniform(8, h - 8), rng.uniform(8, w - 8)
        a, b, th = rng.uniform(3, 9), rng.uniform(2, 6), rng.uniform(0, np.pi)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        sel = (u / a) ** 2 + (v / b) ** 2 <= 1.0
        m[sel & (m == 0)] = l
    return m


# ------------------------------------------------------------------------------------------ CPU: oracle vs OpenCV
@needs_cv2
def test_polygon_oracle_against_opencv():
    tab = synth.make_table(4000, seed=77)
    off, xy = tab.poly_off, tab.poly_xy
    ref = cv2_polygon_features(off, xy)
    got = omorph.polygon_features_csr(off, xy)
    for name in ("area", "perimeter", "centroid_x", "centroid_y", "major_axis_length", "minor_axis_length"):
        np.testing.assert_allclose(got[name], ref[name], rtol=2e-6, err_msg=name)
    np.testing.assert_allclose(got["eccentricity"], ref["eccentricity"], rtol=1e-5, atol=2e-6)


@needs_cv2
def test_polygon_oracle_against_opencv_large_rings():
    # tissue-island sized rings (polygon_morphology.py:240-248 measures these with shapely)
    rng = np.random.default_rng(3)
    rings, off = [], [0]
    for nv in (300, 2500, 20000):
        t = np.sort(rng.uniform(0, 2 * np.pi, nv))
        rad = 800.0 * (1.0 + 0.2 * np.sin(5 * t) + 0.05 * rng.standard_normal(nv))
        rings.append(np.round(np.stack([1000 + rad * np.cos(t), 900 + rad * np.sin(t)], axis=1) * 2) / 2)
        off.append(off[-1] + nv)
    xy = np.concatenate(rings).astype(np.float32)
    off = np.asarray(off, dtype=np.int32)
    ref = cv2_polygon_features(off, xy)
    got = omorph.polygon_features_csr(off, xy)
    for name in ("area", "perimeter", "centroid_x", "centroid_y", "major_axis_length", "minor_axis_length", "eccentricity"):
        np.testing.assert_allclose(got[name], ref[name], rtol=5e-6, err_msg=name)


@needs_cv2
def test_raster_oracle_against_opencv():
    m = blobs(160, 200, 40, seed=5)
    n = int(m.max())
    ref = cv2_raster_props(m, n)
    got = oraster.regionprops(m)
    lab = got["label"] - 1
    assert np.array_equal(got["area"], ref["area"][lab])
    assert np.array_equal(got["bbox"], ref["bbox"][lab])
    np.testing.assert_allclose(got["centroid"][:, 0], ref["centroid_r"][lab], rtol=1e-12)
    np.testing.assert_allclose(got["centroid"][:, 1], ref["centroid_c"][lab], rtol=1e-12)
    for name in ("major_axis_length", "minor_axis_length"):
        np.testing.assert_allclose(got[name], ref[name][lab], rtol=1e-9, err_msg=name)
    np.testing.assert_allclose(got["eccentricity"], ref["eccentricity"][lab], rtol=1e-8, atol=1e-9)


# ------------------------------------------------------------------------------------------ GPU: kernels vs OpenCV
@pytest.mark.gpu
@needs_cv2
@pytest.mark.parametrize("vt", [np.float32, np.float64])
def test_k1_against_opencv(vt):
    from path_gene_multimodal_b200 import map_morph_arrays

    tab = synth.make_table(6000, seed=78, dtype=vt)
    ref = cv2_polygon_features(tab.poly_off, tab.poly_xy)
    res = map_morph_arrays(tab.poly_off, tab.poly_xy, extra=True, write_polygons=False, device=0)
    np.testing.assert_allclose(res["area"], ref["area"], rtol=1e-5)
    np.testing.assert_allclose(res["perimeter"], ref["perimeter"], rtol=1e-5)
    np.testing.assert_allclose(res["centroid_x"], ref["centroid_x"], rtol=1e-7)
    np.testing.assert_allclose(res["centroid_y"], ref["centroid_y"], rtol=1e-7)
    np.testing.assert_allclose(res["major_axis"], ref["major_axis_length"], rtol=1e-5)
    np.testing.assert_allclose(res["minor_axis"], ref["minor_axis_length"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(res["eccentricity"], ref["eccentricity"], rtol=1e-5, atol=2e-6)


@pytest.mark.gpu
@needs_cv2
def test_raster_props_against_opencv():
    from path_gene_multimodal_b200 import raster_regionprops

    m = blobs(256, 320, 90, seed=9)
    ref = cv2_raster_props(m, int(m.max()))
    got = raster_regionprops(m, device=0)
    lab = got["label"].to_numpy() - 1
    assert np.array_equal(lab + 1, np.nonzero(~np.isnan(ref["area"]))[0] + 1)
    assert np.array_equal(got["area"].to_numpy(), ref["area"][lab])
    assert np.array_equal(got[["bbox-0", "bbox-1", "bbox-2", "bbox-3"]].to_numpy(), ref["bbox"][lab])
    np.testing.assert_allclose(got["centroid-0"].to_numpy(), ref["centroid_r"][lab], rtol=1e-12)
    np.testing.assert_allclose(got["centroid-1"].to_numpy(), ref["centroid_c"][lab], rtol=1e-12)
    for name in ("major_axis_length", "minor_axis_length"):
        np.testing.assert_allclose(got[name].to_numpy(), ref[name][lab], rtol=1e-9, err_msg=name)
    np.testing.assert_allclose(got["eccentricity"].to_numpy(), ref["eccentricity"][lab], rtol=1e-8, atol=1e-9)


# ------------------------------------------------------------------------------------------ committed OpenCV outputs
# (oracle/make_cv2_golden.py): the same pins without importing cv2
GOLD = Path(__file__).resolve().parent / "golden" / "opencv_pins.npz"


def _gold_inputs():
    g = np.load(GOLD)
    return g, synth.make_table(1500, seed=79), blobs(192, 224, 60, seed=11)


def test_oracles_against_committed_opencv_vectors():
    g, tab, m = _gold_inputs()
    got = omorph.polygon_features_csr(tab.poly_off, tab.poly_xy)
    for name in ("area", "perimeter", "centroid_x", "centroid_y", "major_axis_length", "minor_axis_length"):
        np.testing.assert_allclose(got[name], g["poly_" + name], rtol=2e-6, err_msg=name)
    np.testing.assert_allclose(got["eccentricity"], g["poly_eccentricity"], rtol=1e-5, atol=2e-6)
    r = oraster.regionprops(m)
    lab = r["label"] - 1
    assert np.array_equal(r["area"], g["raster_area"][lab]) and np.array_equal(r["bbox"], g["raster_bbox"][lab])
    np.testing.assert_allclose(r["centroid"][:, 0], g["raster_centroid_r"][lab], rtol=1e-12)
    np.testing.assert_allclose(r["eccentricity"], g["raster_eccentricity"][lab], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r["major_axis_length"], g["raster_major_axis_length"][lab], rtol=1e-9)
    if cv2 is not None:  # the committed vectors are what this cv2 computes today
        live = cv2_polygon_features(tab.poly_off, tab.poly_xy)
        assert np.array_equal(live["area"], g["poly_area"]) and np.array_equal(live["perimeter"], g["poly_perimeter"])


@pytest.mark.gpu
def test_kernels_against_committed_opencv_vectors():
    from path_gene_multimodal_b200 import map_morph_arrays, raster_regionprops

    g, tab, m = _gold_inputs()
    res = map_morph_arrays(tab.poly_off, tab.poly_xy, extra=True, write_polygons=False, device=0)
    np.testing.assert_allclose(res["area"], g["poly_area"], rtol=1e-5)
    np.testing.assert_allclose(res["perimeter"], g["poly_perimeter"], rtol=1e-5)
    np.testing.assert_allclose(res["centroid_x"], g["poly_centroid_x"], rtol=1e-7)
    np.testing.assert_allclose(res["major_axis"], g["poly_major_axis_length"], rtol=1e-5)
    np.testing.assert_allclose(res["eccentricity"], g["poly_eccentricity"], rtol=1e-5, atol=2e-6)
    got = raster_regionprops(m, device=0)
    lab = got["label"].to_numpy() - 1
    assert np.array_equal(got["area"].to_numpy(), g["raster_area"][lab])
    assert np.array_equal(got[["bbox-0", "bbox-1", "bbox-2", "bbox-3"]].to_numpy(), g["raster_bbox"][lab])
    np.testing.assert_allclose(got["eccentricity"].to_numpy(), g["raster_eccentricity"][lab], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(got["minor_axis_length"].to_numpy(), g["raster_minor_axis_length"][lab], rtol=1e-9)
