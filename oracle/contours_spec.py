"""TEST INFRASTRUCTURE — first-principles restatement of the contour step.  NOT product code.

`cv2.findContours(mask, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)` at
`/root/reference/chips/chips.py:382-386` and what `clean_small_scale_structures` (:405-430) keeps
of it (`contours`, `hierarchy`, `contour_ids` of SURVEY.md §8(a) row H / §8(f) row 4).  OpenCV is a
binary dependency here, so the algorithm is restated from its publication (Suzuki & Abe 1985,
border following with border numbers NBD / LNBD) and pinned by differential tests against the
installed cv2 4.13 (`tests/test_contours_spec.py`): point lists, contour order and hierarchy are
identical on random masks, including thousands of contours per image (border numbers are NOT
wrapped at 127 as in the old 8-bit implementation) and 1-pixel walls (a hole border whose
natural start pixel was marked by an earlier border starts later and is then typed "outer").

What the differential tests establish (the facts a device implementation has to reproduce):
  * raster scan; at pixel p: an OUTER border starts if f(p) == 1 and the west neighbour is 0; else
    a HOLE border starts if f(p) >= 1 and the east neighbour is 0 (a negative mark blocks it);
  * the trace itself never reads the marks: from the start direction (west for outer, east for
    hole) the first non-zero neighbour clockwise, then repeatedly the first non-zero neighbour
    counter-clockwise after the direction we came from; a visited pixel is marked -NBD when its
    east neighbour was examined as a 0-pixel, else +NBD if it was still 1;
  * CHAIN_APPROX_SIMPLE keeps a pixel iff the outgoing direction changes there;
  * parent from LNBD (the last mark met in the row) by Suzuki's table; new contours are inserted
    at the HEAD of their parent's child list and the output is the pre-order walk of that tree:
    siblings appear in reverse order of discovery (bottom of the image first);
  * hierarchy row = [next sibling, previous sibling, first child, parent] in output indices.
This file is only imported by tests.
"""
from __future__ import annotations

import sys

import numpy as np

# direction codes as in OpenCV: 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE  (y grows downwards)
DX = (1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, -1, -1, -1, 0, 1, 1, 1)


def find_contours(mask: np.ndarray, simple: bool = True):
    """Returns (contours: list of int32 [n,1,2] arrays (x, y) in OpenCV's output order,
    hierarchy: int32 [n,4], is_hole: list of bool in the same order)."""
    H, W = mask.shape
    f = np.zeros((H + 2, W + 2), dtype=np.int64)
    f[1:-1, 1:-1] = (np.asarray(mask) != 0)
    nbd = 1
    borders = {1: dict(is_hole=True, parent=0, pts=None)}       # the frame acts as a hole border
    order = []
    for i in range(1, H + 1):
        lnbd = 1
        for j in range(1, W + 1):
            fij = f[i, j]
            if fij == 0:
                continue
            is_hole = None
            if fij == 1 and f[i, j - 1] == 0:
                is_hole, s0 = False, 4
            elif fij >= 1 and f[i, j + 1] == 0:
                is_hole, s0 = True, 0
                if fij > 1:
                    lnbd = fij
            if is_hole is not None:
                nbd += 1
                b0 = borders[lnbd]
                if is_hole:
                    parent = lnbd if not b0["is_hole"] else b0["parent"]
                else:
                    parent = lnbd if b0["is_hole"] else b0["parent"]
                pts = []
                s = s0
                found = False
                for _ in range(8):                               # clockwise from the start direction
                    s = (s - 1) & 7
                    if f[i + DY[s], j + DX[s]] != 0:
                        found = True
                        break
                if not found:                                    # isolated pixel
                    f[i, j] = -nbd
                    pts.append((j - 1, i - 1))
                else:
                    i1, j1 = i + DY[s], j + DX[s]
                    i3, j3 = i, j
                    prev_s = s ^ 4
                    while True:
                        s_end = s
                        while True:                              # counter-clockwise search
                            s += 1
                            i4, j4 = i3 + DY[s & 7], j3 + DX[s & 7]
                            if f[i4, j4] != 0:
                                break
                        s &= 7
                        if 0 <= s - 1 < s_end:                   # the east neighbour was examined as 0
                            f[i3, j3] = -nbd
                        elif f[i3, j3] == 1:
                            f[i3, j3] = nbd
                        if (not simple) or s != prev_s:
                            pts.append((j3 - 1, i3 - 1))
                            prev_s = s
                        if (i4, j4) == (i, j) and (i3, j3) == (i1, j1):
                            break
                        i3, j3 = i4, j4
                        s = (s + 4) & 7
                borders[nbd] = dict(is_hole=is_hole, parent=parent,
                                    pts=np.array(pts, dtype=np.int32).reshape(-1, 1, 2))
                order.append(nbd)
            if f[i, j] != 1:
                lnbd = abs(f[i, j])
    children = {}
    for b in order:
        children.setdefault(borders[b]["parent"], []).insert(0, b)
    out = []
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 20000))

    def walk(p):
        for c in children.get(p, []):
            out.append(c)
            walk(c)
    try:
        walk(1)
    finally:
        sys.setrecursionlimit(old)
    idx = {b: k for k, b in enumerate(out)}
    hier = np.full((len(out), 4), -1, dtype=np.int32)
    for p, ch in children.items():
        for k, c in enumerate(ch):
            if k + 1 < len(ch):
                hier[idx[c], 0] = idx[ch[k + 1]]
            if k > 0:
                hier[idx[c], 1] = idx[ch[k - 1]]
            hier[idx[c], 3] = idx[p] if p in idx else -1
        if p in idx and ch:
            hier[idx[p], 2] = idx[ch[0]]
    return [borders[b]["pts"] for b in out], hier, [borders[b]["is_hole"] for b in out]


def contour_area(cnt: np.ndarray) -> float:
    """cv2.contourArea: |shoelace| / 2 over the polygon through the contour's points."""
    p = cnt[:, 0, :].astype(np.float64)
    x, y = p[:, 0], p[:, 1]
    return float(abs(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)) * 0.5)


def clean_small_scale_structures(contours, hierarchy, shape, disk_area, area_threshold):
    """chips.py:405-430 on the restated contours: (kept contours, their hierarchy rows, filled map).
    The fill of a kept top-level contour is its outer polygon: even-odd rule on pixel centres,
    boundary pixels included (what cv2.drawContours(..., -1) paints)."""
    import cv2      # only for the polygon fill of this checker
    dxmap = np.zeros(shape, dtype=np.float64)
    new_contours, new_hierarchy = [], []
    for i, cnt in enumerate(contours):
        if contour_area(cnt) / disk_area >= area_threshold:
            new_contours.append(cnt)
            if hierarchy[i][3] < 0:
                cv2.drawContours(dxmap, [cnt], -1, (255, 255, 255), -1)
            new_hierarchy.append(hierarchy[i])
    return new_contours, np.array(new_hierarchy), dxmap / 255
