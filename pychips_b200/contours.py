"""Host side of the device contour pass.

`chips_contours` (csrc/contours.cu) returns one record per border in ORDER OF DISCOVERY
(border number NBD = index + 2, Suzuki-Abe) with its parent border number, and the kept points of
every border.  What is left for the host is the O(#borders) bookkeeping of
`cv2.findContours(..., RETR_TREE, ...)`'s output order — children of a border are listed in
reverse order of discovery, the tree in pre-order — and the `[next, prev, first_child, parent]`
hierarchy rows (/root/reference/chips/chips.py:382-386 consumes both).
"""
from __future__ import annotations

import numpy as np

REC_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("is_hole", "<i4"), ("parent", "<i4"),
                      ("n_points", "<i4"), ("offset", "<i4"), ("area2", "<i8")])
COUNTS_DTYPE = np.dtype([("n_border_px", "<i4"), ("n_contours", "<i4"), ("n_points", "<i4"),
                         ("changed", "<i4"), ("overflow", "<i4"), ("pad_", "<i4")])


def output_order(parent_nbd: np.ndarray):
    """parent_nbd[b]: border number of the parent of border b (1 = the frame).  Returns
    (seq, hierarchy): seq[i] = discovery index of the i-th output contour, hierarchy int32 [n, 4]."""
    n = len(parent_nbd)
    hier = np.full((n, 4), -1, dtype=np.int32)
    if n == 0:
        return np.zeros(0, dtype=np.int64), hier
    parent = np.asarray(parent_nbd, dtype=np.int64) - 2          # -1 = frame
    if (parent < -1).any():
        raise RuntimeError("contour pass returned a border without a parent")
    b = np.arange(n)
    by_parent = np.lexsort((-b, parent))                          # siblings: latest discovered first
    first_child = np.full(n + 1, -1, dtype=np.int64)              # slot n = frame
    next_sib = np.full(n, -1, dtype=np.int64)
    prev_sib = np.full(n, -1, dtype=np.int64)
    ps = parent[by_parent]
    head = np.ones(n, dtype=bool)
    head[1:] = ps[1:] != ps[:-1]
    first_child[np.where(ps[head] < 0, n, ps[head])] = by_parent[head]
    same = ~head[1:]
    next_sib[by_parent[:-1][same]] = by_parent[1:][same]
    prev_sib[by_parent[1:][same]] = by_parent[:-1][same]
    seq = _preorder_by_levels(parent, by_parent, head)
    if seq is None:                                               # very deep nesting: plain walk
        seq = _preorder_walk(first_child, next_sib, n)
    pos = np.empty(n, dtype=np.int64)
    pos[seq] = np.arange(n)
    look = lambda a: np.where(a >= 0, pos[np.maximum(a, 0)], -1)
    hier[:, 0] = look(next_sib[seq])
    hier[:, 1] = look(prev_sib[seq])
    hier[:, 2] = look(first_child[:n][seq])
    hier[:, 3] = look(parent[seq])
    return seq, hier


def _preorder_walk(first_child, next_sib, n):
    seq = np.empty(n, dtype=np.int64)
    k = 0
    stack = [int(first_child[n])]
    fc, ns = first_child.tolist(), next_sib.tolist()
    while stack:
        node = stack.pop()
        if node < 0:
            continue
        seq[k] = node
        k += 1
        stack.append(ns[node])
        stack.append(fc[node])
    if k != n:
        raise RuntimeError("contour tree is not connected to the frame")
    return seq


def _preorder_by_levels(parent, by_parent, head, max_depth: int = 48):
    """Pre-order positions without walking the tree: position = parent's position + 1 + sizes of
    the subtrees of the siblings listed before it.  O(depth) numpy passes; None if deeper than
    `max_depth` (concentric rings), where the plain walk is used."""
    n = len(parent)
    levels = [np.flatnonzero(parent < 0)]
    seen = len(levels[0])
    while seen < n:
        if len(levels) > max_depth or len(levels[-1]) == 0:
            return None if len(levels) > max_depth else _fail()
        mark = np.zeros(n, dtype=bool)
        mark[levels[-1]] = True
        nxt = np.flatnonzero((parent >= 0) & mark[np.maximum(parent, 0)])
        levels.append(nxt)
        seen += len(nxt)
    size = np.ones(n, dtype=np.int64)
    for lev in reversed(levels[1:]):
        np.add.at(size, parent[lev], size[lev])
    # exclusive sums of the subtree sizes inside every sibling group, in sibling order
    sz = size[by_parent]
    run = np.cumsum(sz) - sz
    group_start = np.maximum.accumulate(np.where(head, np.arange(n), 0))
    before = np.empty(n, dtype=np.int64)
    before[by_parent] = run - run[group_start]
    pos = np.empty(n, dtype=np.int64)
    pos[levels[0]] = before[levels[0]]
    for lev in levels[1:]:
        pos[lev] = pos[parent[lev]] + 1 + before[lev]
    seq = np.empty(n, dtype=np.int64)
    seq[pos] = np.arange(n)
    return seq


def _fail():
    raise RuntimeError("contour tree is not connected to the frame")


def assemble(recs: np.ndarray, points: np.ndarray):
    """Records + point buffer -> (contours, hierarchy, twice_area) exactly as
    cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE) orders them; twice_area[i] =
    2 * cv2.contourArea(contours[i]) as an exact integer."""
    seq, hier = output_order(recs["parent"])
    pts = points.reshape(-1, 1, 2)                   # OpenCV's [n, 1, 2] layout; the slices are views
    start = recs["offset"][seq].astype(np.int64)
    contours = [pts[o:e] for o, e in zip(start.tolist(), (start + recs["n_points"][seq]).tolist())]
    return contours, hier, np.abs(recs["area2"][seq])
