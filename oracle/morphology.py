"""Oracle (test infrastructure): per-polygon morphology in float64.

Restates, because shapely / GEOS and scikit-image are not importable here (un-vendored,
un-pinned third-party dependencies of the reference):

* shapely ``Polygon.area`` / ``.length`` / ``.centroid`` / ``.bounds`` as called at
  /root/reference/polygon_morphology.py:240-248 and
  /root/reference/create_and_overlay_polygon_from_prediction.py:298-299.
  GEOS algorithms (published): ``Area::ofRingSigned`` = 1/2 * sum (x_i - x_0) (y_{i-1} - y_{i+1}),
  ``Length::ofLine`` = sum of segment lengths over the closed ring, area-weighted centroid, min/max.
* the cell-18 feature definitions of hovernet_tile_inference.ipynb:2415-2456
  (skimage ``eccentricity = sqrt(1 - l2/l1)``, ``major = 4 sqrt(l1)``, ``minor = 4 sqrt(l2)`` from the
  second central moments; derived ``perimeter_area, compactness, roundness, elongation``),
  evaluated on the polygon's exact area moments instead of a raster.

Pinning: derived-feature formulas are pinned on the ten stored notebook rows (Appendix D-2,
tests/golden/notebook_known_answers.json); area / perimeter / centroid / moments are checked on
analytic shapes.  Polygon-vs-raster eccentricity is "parity unpinned" by construction.
"""
from __future__ import annotations

import numpy as np

FEATURES = ["area", "perimeter", "eccentricity", "circularity", "centroid_x", "centroid_y",
            "major_axis_length", "minor_axis_length",
            "bbox_xmin", "bbox_ymin", "bbox_xmax", "bbox_ymax"]


def polygon_features_one(poly):
    """Scalar-loop version for one ring (list of [x, y]; closing vertex optional)."""
    pts = [(float(x), float(y)) for x, y in poly]
    n = len(pts)
    nan = float("nan")
    if n < 3:
        return dict.fromkeys(FEATURES, nan)
    x0, y0 = pts[0]
    s2 = sx = sy = ixx = iyy = ixy = per = 0.0
    for i in range(n):
        xa, ya = pts[i][0] - x0, pts[i][1] - y0
        xb, yb = pts[(i + 1) % n][0] - x0, pts[(i + 1) % n][1] - y0
        a = xa * yb - xb * ya
        s2 += a
        sx += (xa + xb) * a
        sy += (ya + yb) * a
        ixx += (ya * ya + ya * yb + yb * yb) * a
        iyy += (xa * xa + xa * xb + xb * xb) * a
        ixy += (xa * yb + 2.0 * xa * ya + 2.0 * xb * yb + xb * ya) * a
        per += np.sqrt((xb - xa) ** 2 + (yb - ya) ** 2)
    xs = [p[0] for p in pts]
    ys = [p[1] for p in pts]
    out = {"perimeter": per, "bbox_xmin": min(xs), "bbox_ymin": min(ys), "bbox_xmax": max(xs), "bbox_ymax": max(ys)}
    A = 0.5 * s2
    out["area"] = abs(A)
    out["circularity"] = 4.0 * np.pi * abs(A) / max(per, 1.0) ** 2
    if A == 0.0:
        for k in ("eccentricity", "centroid_x", "centroid_y", "major_axis_length", "minor_axis_length"):
            out[k] = nan
        return out
    cx, cy = sx / (6.0 * A), sy / (6.0 * A)
    mu20 = iyy / (12.0 * A) - cx * cx
    mu02 = ixx / (12.0 * A) - cy * cy
    mu11 = ixy / (24.0 * A) - cx * cy
    m = 0.5 * (mu20 + mu02)
    c = np.sqrt((0.5 * (mu20 - mu02)) ** 2 + mu11 * mu11)
    l1, l2 = m + c, max(m - c, 0.0)
    out["eccentricity"] = float(np.sqrt(1.0 - l2 / l1)) if l1 > 0 else 0.0
    out["centroid_x"], out["centroid_y"] = cx + x0, cy + y0
    out["major_axis_length"] = 4.0 * np.sqrt(max(l1, 0.0))
    out["minor_axis_length"] = 4.0 * np.sqrt(l2)
    return out


def polygon_features_csr(poly_off, poly_xy):
    """Vectorised float64 features over CSR rings. Returns dict name -> float64[N].

    Rows with fewer than 3 vertices give NaN everywhere (shapely would raise); zero-area rings
    keep perimeter / bbox / area=0 / circularity=0 and NaN for centroid, axes, eccentricity.
    """
    off = np.asarray(poly_off, dtype=np.int64)
    xy = np.asarray(poly_xy, dtype=np.float64)
    n = len(off) - 1
    nv = np.diff(off)
    m = int(off[-1])
    owner = np.repeat(np.arange(n), nv)
    idx = np.arange(m)
    nxt = idx + 1
    last = off[1:][nv > 0] - 1
    nxt[last] = off[:-1][nv > 0]
    first = off[:-1][owner]
    xa = xy[:, 0] - xy[first, 0]
    ya = xy[:, 1] - xy[first, 1]
    xb, yb = xa[nxt], ya[nxt]
    a = xa * yb - xb * ya

    def seg(v):
        out = np.zeros(n, dtype=np.float64)
        np.add.at(out, owner, v)  # sequential per-row accumulation, same order as the scalar loop
        return out

    s2 = seg(a)
    sx = seg((xa + xb) * a)
    sy = seg((ya + yb) * a)
    ixx = seg((ya * ya + ya * yb + yb * yb) * a)
    iyy = seg((xa * xa + xa * xb + xb * xb) * a)
    ixy = seg((xa * yb + 2.0 * xa * ya + 2.0 * xb * yb + xb * ya) * a)
    per = seg(np.sqrt((xb - xa) ** 2 + (yb - ya) ** 2))
    nan = np.full(n, np.nan)
    xmin, ymin, xmax, ymax = nan.copy(), nan.copy(), nan.copy(), nan.copy()
    nz = nv > 0
    if m:
        xmin[nz] = np.minimum.reduceat(xy[:, 0], off[:-1][nz])
        xmax[nz] = np.maximum.reduceat(xy[:, 0], off[:-1][nz])
        ymin[nz] = np.minimum.reduceat(xy[:, 1], off[:-1][nz])
        ymax[nz] = np.maximum.reduceat(xy[:, 1], off[:-1][nz])
    A = 0.5 * s2
    ok = (nv >= 3)
    good = ok & (A != 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        cx, cy = sx / (6.0 * A), sy / (6.0 * A)
        mu20 = iyy / (12.0 * A) - cx * cx
        mu02 = ixx / (12.0 * A) - cy * cy
        mu11 = ixy / (24.0 * A) - cx * cy
        mm = 0.5 * (mu20 + mu02)
        cc = np.sqrt((0.5 * (mu20 - mu02)) ** 2 + mu11 * mu11)
        l1 = mm + cc
        l2 = np.maximum(mm - cc, 0.0)
        ecc = np.where(l1 > 0, np.sqrt(1.0 - l2 / l1), 0.0)
        major = 4.0 * np.sqrt(np.maximum(l1, 0.0))
        minor = 4.0 * np.sqrt(l2)
    x0 = np.zeros(n)
    y0 = np.zeros(n)
    x0[nz] = xy[off[:-1][nz], 0]
    y0[nz] = xy[off[:-1][nz], 1]
    out = {
        "area": np.where(ok, np.abs(A), np.nan),
        "perimeter": np.where(ok, per, np.nan),
        "eccentricity": np.where(good, ecc, np.nan),
        "circularity": np.where(ok, 4.0 * np.pi * np.abs(A) / np.maximum(per, 1.0) ** 2, np.nan),
        "centroid_x": np.where(good, cx + x0, np.nan),
        "centroid_y": np.where(good, cy + y0, np.nan),
        "major_axis_length": np.where(good, major, np.nan),
        "minor_axis_length": np.where(good, minor, np.nan),
        "bbox_xmin": np.where(ok, xmin, np.nan), "bbox_ymin": np.where(ok, ymin, np.nan),
        "bbox_xmax": np.where(ok, xmax, np.nan), "bbox_ymax": np.where(ok, ymax, np.nan),
    }
    return out


def geos_area_length(poly):
    """GEOS-order area / length for one ring (closing vertex added if absent)."""
    pts = [(float(x), float(y)) for x, y in poly]
    if pts[0] != pts[-1]:
        pts.append(pts[0])
    n = len(pts)
    x0 = pts[0][0]
    s = 0.0
    for i in range(1, n - 1):
        s += (pts[i][0] - x0) * (pts[i - 1][1] - pts[i + 1][1])
    length = 0.0
    for i in range(n - 1):
        length += np.sqrt((pts[i + 1][0] - pts[i][0]) ** 2 + (pts[i + 1][1] - pts[i][1]) ** 2)
    return abs(s / 2.0), length


def derived_features(area, perimeter, major, minor):
    """Cell-18 derived columns (ipynb:2431-2456), ``clip(lower=1)`` guards included."""
    area = np.asarray(area, dtype=np.float64)
    perimeter = np.asarray(perimeter, dtype=np.float64)
    major = np.asarray(major, dtype=np.float64)
    minor = np.asarray(minor, dtype=np.float64)
    return {
        "perimeter_area": perimeter / np.maximum(area, 1.0),
        "compactness": 4.0 * np.pi * area / np.maximum(perimeter, 1.0) ** 2,
        "roundness": 4.0 * area / (np.pi * np.maximum(major, 1.0) ** 2),
        "elongation": major / np.maximum(minor, 1.0),
        "eccentricity": np.sqrt(1.0 - (minor / major) ** 2),
    }


def zscore(col):
    """Cell 21 (ipynb:2903): (x - mean) / std(ddof=0); all zeros when sigma is 0 or NaN."""
    col = np.asarray(col, dtype=np.float64)
    mu = np.nanmean(col) if col.size else np.nan
    sigma = np.nanstd(col) if col.size else np.nan
    if sigma == 0 or np.isnan(sigma):
        return np.zeros_like(col)
    return (col - mu) / sigma
