"""TEST INFRASTRUCTURE ONLY - CPU oracle for the instance-map -> polygon step (second half).

Reference: /root/reference/aggregated_hovernet_run.py:183-198 - per instance
    mask = (inst_map == inst_id); contours = find_contours(mask.astype(float), level=0.5)
    contour = max(contours, key=lambda c: c.shape[0]); poly = np.stack([xs, ys], 1)
    poly = approximate_polygon(poly, tolerance=0.5).tolist()
skimage is NOT installed in this image, so these are restatements of the published skimage.measure algorithms
(v0.19-0.25: _find_contours_cy.pyx `_get_contour_segments`, _find_contours.py `_assemble_contours`,
_polygon.py `approximate_polygon`), statement for statement where order matters.  PARITY UNPINNED by reference
fixtures (no stored polygons whose instance map is recoverable).

`approximate_polygon` keeps a vertex iff its distance to the chord is `> tolerance`, evaluated through
arctan2 / sin / cos.  On the half-pixel lattice of a binary mask's contour with tolerance 0.5 the distance can be
EXACTLY 0.5 (axis-parallel chords - there the trigonometric values are exact and the float result is 0.5, not kept
- and chords whose doubled components form a Pythagorean triple, where the float result is 0.5 +- 1 ulp depending
on libm).  `approximate_polygon_exact` evaluates the same decisions in integer arithmetic on the doubled
coordinates (a tie is "not greater", as in exact mathematics); `approximate_polygon_float` is skimage's float code,
kept to count how often the two differ.  The product implements the exact rule.
"""
from collections import deque

import numpy as np


def contour_segments(mask):
    """Directed segments ((r, c) -> (r, c)) of the level-0.5 iso-line of a 0/1 array, in skimage's square order
    (row-major over the 2x2 squares) and with its per-case directions (fully_connected='low')."""
    m = np.asarray(mask).astype(bool)
    h, w = m.shape
    segs = []
    for r0 in range(h - 1):
        for c0 in range(w - 1):
            r1, c1 = r0 + 1, c0 + 1
            ul, ur, ll, lr = m[r0, c0], m[r0, c1], m[r1, c0], m[r1, c1]
            case = int(ul) + 2 * int(ur) + 4 * int(ll) + 8 * int(lr)
            if case in (0, 15):
                continue
            top, bottom = (r0, c0 + 0.5), (r1, c0 + 0.5)
            left, right = (r0 + 0.5, c0), (r0 + 0.5, c1)
            if case == 1: segs.append((top, left))
            elif case == 2: segs.append((right, top))
            elif case == 3: segs.append((right, left))
            elif case == 4: segs.append((left, bottom))
            elif case == 5: segs.append((top, bottom))
            elif case == 6: segs.append((right, top)); segs.append((left, bottom))
            elif case == 7: segs.append((right, bottom))
            elif case == 8: segs.append((bottom, right))
            elif case == 9: segs.append((top, left)); segs.append((bottom, right))
            elif case == 10: segs.append((bottom, top))
            elif case == 11: segs.append((bottom, left))
            elif case == 12: segs.append((left, right))
            elif case == 13: segs.append((top, right))
            elif case == 14: segs.append((left, top))
    return segs


def assemble_contours(segments):
    """skimage.measure._find_contours._assemble_contours, literally."""
    current_index = 0
    contours = {}
    starts = {}
    ends = {}
    for from_point, to_point in segments:
        if from_point == to_point:
            continue
        tail, tail_num = starts.pop(to_point, (None, None))
        head, head_num = ends.pop(from_point, (None, None))
        if tail is not None and head is not None:
            if tail is head:
                head.append(to_point)
            elif tail_num > head_num:
                head.extend(tail)
                contours.pop(tail_num, None)
                starts[head[0]] = (head, head_num)
                ends[head[-1]] = (head, head_num)
            else:
                tail.extendleft(reversed(head))
                starts.pop(head[0], None)
                contours.pop(head_num, None)
                starts[tail[0]] = (tail, tail_num)
                ends[tail[-1]] = (tail, tail_num)
        elif tail is None and head is None:
            new_contour = deque((from_point, to_point))
            contours[current_index] = new_contour
            starts[from_point] = (new_contour, current_index)
            ends[to_point] = (new_contour, current_index)
            current_index += 1
        elif head is None:
            tail.appendleft(from_point)
            starts[from_point] = (tail, tail_num)
        else:
            head.append(to_point)
            ends[to_point] = (head, head_num)
    return [np.array(contour) for _, contour in sorted(contours.items())]


def find_contours(mask):
    """find_contours(mask.astype(float), level=0.5): list of (n, 2) arrays of (row, col)."""
    return assemble_contours(contour_segments(mask))


def approximate_polygon_float(coords, tolerance):
    """skimage.measure.approximate_polygon, literally (float trigonometry)."""
    coords = np.asarray(coords, dtype=np.float64)
    if tolerance <= 0:
        return coords
    chain = np.zeros(coords.shape[0], "bool")
    dists = np.zeros(coords.shape[0])
    chain[0] = True
    chain[-1] = True
    pos_stack = [(0, chain.shape[0] - 1)]
    end_of_chain = False
    while not end_of_chain:
        start, end = pos_stack.pop()
        r0, c0 = coords[start, :]
        r1, c1 = coords[end, :]
        dr = r1 - r0
        dc = c1 - c0
        segment_angle = -np.arctan2(dr, dc)
        segment_dist = c0 * np.sin(segment_angle) + r0 * np.cos(segment_angle)
        segment_coords = coords[start + 1:end, :]
        segment_dists = dists[start + 1:end]
        dr0 = segment_coords[:, 0] - r0
        dc0 = segment_coords[:, 1] - c0
        dr1 = segment_coords[:, 0] - r1
        dc1 = segment_coords[:, 1] - c1
        projected_lengths0 = dr0 * dr + dc0 * dc
        projected_lengths1 = -dr1 * dr - dc1 * dc
        perp = np.logical_and(projected_lengths0 > 0, projected_lengths1 > 0)
        eucl = np.logical_not(perp)
        segment_dists[perp] = np.abs(segment_coords[perp, 0] * np.cos(segment_angle)
                                     + segment_coords[perp, 1] * np.sin(segment_angle) - segment_dist)
        segment_dists[eucl] = np.minimum(np.sqrt(dc0[eucl] ** 2 + dr0[eucl] ** 2), np.sqrt(dc1[eucl] ** 2 + dr1[eucl] ** 2))
        if np.any(segment_dists > tolerance):
            new_end = start + np.argmax(segment_dists) + 1
            pos_stack.append((new_end, end))
            pos_stack.append((start, new_end))
            chain[new_end] = True
        if len(pos_stack) == 0:
            end_of_chain = True
    return coords[chain, :]


def approximate_polygon_exact(coords, tolerance=0.5):
    """The same recursion with every decision in integer arithmetic. coords are multiples of 0.5; tolerance is a
    multiple of 0.5. Squared distance of point p to chord (a, b) as a fraction num / den:
      perpendicular case (both projections > 0): cross(b - a, p - a)^2 / |b - a|^2
      otherwise: min(|p - a|^2, |p - b|^2) / 1.
    keep iff distance > tolerance; the first point of maximum distance splits (np.argmax)."""
    pts = np.rint(np.asarray(coords, dtype=np.float64) * 2).astype(np.int64)   # doubled coordinates
    n = len(pts)
    tol2 = int(round(tolerance * 2)) ** 2                                     # squared, in doubled units
    chain = np.zeros(n, dtype=bool)
    chain[0] = chain[-1] = True
    stack = [(0, n - 1)]
    while stack:
        start, end = stack.pop()
        a, b = pts[start], pts[end]
        d = b - a
        L2 = int(d[0] * d[0] + d[1] * d[1])
        best_num, best_den, best_i = 0, 1, -1
        for i in range(start + 1, end):
            p = pts[i]
            pa, pb = p - a, p - b
            proj0 = int(pa[0] * d[0] + pa[1] * d[1])
            proj1 = int(-(pb[0] * d[0] + pb[1] * d[1]))
            if proj0 > 0 and proj1 > 0:
                cr = int(d[0] * pa[1] - d[1] * pa[0])
                num, den = cr * cr, L2
            else:
                num, den = min(int(pa[0] * pa[0] + pa[1] * pa[1]), int(pb[0] * pb[0] + pb[1] * pb[1])), 1
            if best_i < 0 or num * best_den > best_num * den:                 # strictly greater: first maximum wins
                best_num, best_den, best_i = num, den, i
        if best_i >= 0 and best_num > tol2 * best_den:
            stack.append((best_i, end))
            stack.append((start, best_i))
            chain[best_i] = True
    return np.asarray(coords, dtype=np.float64)[chain]


def instance_polygons(inst_map, tolerance=0.5, exact=True):
    """poly_dict of aggregated_hovernet_run.py:183-198: label -> list of [x, y] (closed ring: first == last)."""
    m = np.asarray(inst_map)
    out = {}
    approx = approximate_polygon_exact if exact else approximate_polygon_float
    for lab in [int(v) for v in np.unique(m) if v > 0]:
        # the reference masks the whole tile; squares without a pixel of the label yield no segment, so the mask
        # cropped to the bounding box + 1 pixel gives the same segments in the same order (and finishes in time)
        rr, cc = np.nonzero(m == lab)
        r0, c0 = max(rr.min() - 1, 0), max(cc.min() - 1, 0)
        r1, c1 = min(rr.max() + 2, m.shape[0]), min(cc.max() + 2, m.shape[1])
        contours = [c + np.array([r0, c0], dtype=np.float64) for c in find_contours(m[r0:r1, c0:c1] == lab)]
        if not contours:
            continue
        contour = max(contours, key=lambda c: c.shape[0])
        poly = np.stack([contour[:, 1], contour[:, 0]], axis=1)               # (x, y) = (col, row)
        out[lab] = approx(poly, tolerance).tolist()
    return out


def longest_contour_by_cycles(mask):
    """The rule the CUDA kernel uses instead of the dictionary bookkeeping (equivalent; checked in tests): the
    segments of one label form disjoint directed cycles (open chains only at the image border); a cycle leaves
    _assemble_contours starting at the to-point of its LAST segment in square order (the segment that closes it),
    contours are numbered by their FIRST segment, and max(..., key=len) takes the first longest."""
    segs = contour_segments(mask)
    if not segs:
        return None
    nxt = {}
    for i, (f, _) in enumerate(segs):
        nxt[f] = i
    has_pred = set(t for _, t in segs)
    seen = [False] * len(segs)
    best = None
    order = list(range(len(segs)))
    for i in order:
        if seen[i]:
            continue
        # an open chain must be entered at its first segment; chains are found when i has no predecessor,
        # otherwise i lies on a cycle or in the middle of a chain that is visited from its own start
        f, _ = segs[i]
        j = i
        if f in has_pred:
            # walk forward to see whether we return to i (cycle) or fall off the border (chain, handled from its start)
            k, steps, closed = i, 0, False
            while True:
                k2 = nxt.get(segs[k][1])
                steps += 1
                if k2 is None:
                    break
                if k2 == i:
                    closed = True
                    break
                k = k2
            if not closed:
                continue
        idxs = []
        k = j
        while k is not None and not seen[k]:
            seen[k] = True
            idxs.append(k)
            k = nxt.get(segs[k][1])
        closed = k is not None and k == j
        if closed:
            last = max(idxs)
            pos = idxs.index(last)
            ordered = idxs[pos + 1:] + idxs[:pos + 1]                          # starts right after the closing segment
            pts = [segs[ordered[0]][0]] + [segs[q][1] for q in ordered]
        else:
            pts = [segs[idxs[0]][0]] + [segs[q][1] for q in idxs]
        first = min(idxs)
        key = (-len(pts), first)
        if best is None or key < best[0]:
            best = (key, np.array(pts))
    return best[1]
