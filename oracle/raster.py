"""TEST INFRASTRUCTURE ONLY - CPU oracle for the raster region properties.

skimage is the reference's dependency here (regionprops / regionprops_table, aggregated_hovernet_run.py:172,
hovernet_tile_inference.ipynb:2415) and it is NOT installed in this image, so this is a restatement of the
published skimage.measure algorithms (v0.19-0.25: _regionprops.py, _moments.py, _regionprops_utils.perimeter),
region by region like skimage does, using scipy.ndimage for the erosion / convolution exactly as skimage's
perimeter() does.  PARITY UNPINNED by reference fixtures: the notebook's stored rows (Appendix D-2) have no
recoverable inst_map; they pin only the derived-feature formulas.  Analytic shapes pin the restatement.
"""
import numpy as np
from scipy import ndimage as ndi


def _perimeter(image):
    """skimage.measure.perimeter(image, neighborhood=4)."""
    strel = ndi.generate_binary_structure(2, 1)
    eroded = ndi.binary_erosion(image, strel, border_value=0)
    border = image.astype(np.uint8) - eroded.astype(np.uint8)
    weights = np.zeros(50, dtype=np.float64)
    weights[[5, 7, 15, 17, 25, 27]] = 1
    weights[[21, 33]] = np.sqrt(2)
    weights[[13, 23]] = (1 + np.sqrt(2)) / 2
    conv = ndi.convolve(border, np.array([[10, 2, 10], [2, 1, 2], [10, 2, 10]]), mode="constant", cval=0)
    hist = np.bincount(conv.ravel(), minlength=50)
    return float(hist @ weights)


def _convex_area(img):
    """np.sum(skimage.morphology.convex_hull_image(img)): hull (qhull) of the pixels' diamond corners
    (r +- 1/2, c), (r, c +- 1/2), then every pixel centre inside the hull or on its border (grid_points_in_poly
    with include_borders=True). The inside test is done here in integers on doubled coordinates."""
    from scipy.spatial import ConvexHull

    rr, cc = np.nonzero(img)
    pts = np.concatenate([np.stack([2 * rr + dr, 2 * cc + dc], axis=1) for dr, dc in ((-1, 0), (1, 0), (0, -1), (0, 1))])
    pts = np.unique(pts, axis=0)
    hull = ConvexHull(pts.astype(np.float64))
    v = pts[hull.vertices].astype(np.int64)          # counter-clockwise
    if len(v) >= 3:
        a = v
        b = np.roll(v, -1, axis=0)
        area2 = np.sum(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0])
        if area2 < 0:
            v = v[::-1]
    h, w = img.shape
    gr, gc = np.mgrid[0:h, 0:w]
    p = np.stack([2 * gr.ravel(), 2 * gc.ravel()], axis=1).astype(np.int64)
    inside = np.ones(len(p), dtype=bool)
    for i in range(len(v)):
        a, b = v[i], v[(i + 1) % len(v)]
        cross = (b[0] - a[0]) * (p[:, 1] - a[1]) - (b[1] - a[1]) * (p[:, 0] - a[0])
        inside &= cross >= 0
    return int(inside.sum())


def regionprops(inst_map):
    """dict of arrays, one entry per label present (ascending)."""
    m = np.asarray(inst_map)
    labels = [int(l) for l in np.unique(m) if l > 0]
    out = {k: [] for k in ("label", "area", "bbox", "centroid", "perimeter", "eccentricity", "major_axis_length",
                           "minor_axis_length", "orientation", "solidity")}
    for l in labels:
        rows, cols = np.nonzero(m == l)
        r0, r1, c0, c1 = rows.min(), rows.max() + 1, cols.min(), cols.max() + 1
        img = (m[r0:r1, c0:c1] == l)
        rr, cc = np.nonzero(img)
        area = float(len(rr))
        cr, ccn = rr.mean(), cc.mean()
        dr, dc = rr - cr, cc - ccn
        mu20, mu02, mu11 = (dr * dr).sum(), (dc * dc).sum(), (dr * dc).sum()
        t = np.array([[mu02, -mu11], [-mu11, mu20]]) / area          # inertia_tensor
        ev = np.clip(np.sort(np.linalg.eigvalsh(t))[::-1], 0, None)   # inertia_tensor_eigvals
        l1, l2 = ev
        a, b, c = t[0, 0], t[0, 1], t[1, 1]
        if a - c == 0:
            orient = np.pi / 4 if b < 0 else -np.pi / 4
        else:
            orient = 0.5 * np.arctan2(-2 * b, c - a)
        out["label"].append(l)
        out["area"].append(area)
        out["bbox"].append([r0, c0, r1, c1])
        out["centroid"].append([cr + r0, ccn + c0])
        out["perimeter"].append(_perimeter(img))
        out["eccentricity"].append(0.0 if l1 == 0 else float(np.sqrt(1 - l2 / l1)))
        out["major_axis_length"].append(4 * np.sqrt(l1))
        out["minor_axis_length"].append(4 * np.sqrt(l2))
        out["orientation"].append(orient)
        out["solidity"].append(area / _convex_area(img))
    return {k: np.array(v) for k, v in out.items()}
