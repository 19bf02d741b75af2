"""TEST INFRASTRUCTURE — executable spec of the PARALLEL formulation of the contour step.  NOT product code.

`oracle/contours_spec.py` restates Suzuki-Abe border following as published (one raster scan that
traces and marks as it goes).  This file restates the same result in the form a device
implementation needs, and `tests/test_contours_spec.py` checks it against OpenCV:

  (i)   states (p, s) = "at foreground pixel p, having arrived from its neighbour in direction s";
        the border-following step is a PERMUTATION of the states (first non-zero neighbour
        counter-clockwise after s), and per state, locally: the next state, whether the visit
        marks p negative (east neighbour examined as 0), whether CHAIN_APPROX_SIMPLE keeps p
        (outgoing direction != travel direction);
  (ii)  borders = cycles of that permutation (pointer doubling, label = smallest state id);
        per pixel the cycles through it and whether each marks it negative;
  (iii) the only sequential part: candidate pixels (west or east neighbour 0) in raster order —
        f(p) "now" follows from the cycles through p that have already started (negative mark of
        the latest such cycle, else positive mark of the earliest, else 1); an outer border starts
        at f == 1 with west 0, else a hole border at f >= 1 with east 0; border number = order of
        start; parent from the last mark met in the row (Suzuki's table);
  (iv)  output: children in reverse order of discovery, pre-order; the points of a border are the
        kept states along its cycle from the start state.
No trace ever reads a mark, so (i), (ii) and (iv) are data parallel; (iii) touches only border
pixels and moves a start away from a cycle's first candidate only at pixels shared with a border
that started earlier (1-pixel walls).
"""
from __future__ import annotations

import sys

import numpy as np
DX = (1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, -1, -1, -1, 0, 1, 1, 1)
DX_arr, DY_arr = np.array(DX), np.array(DY)

def find_contours_parallel(mask, simple=True):
    H, W = mask.shape
    fg = np.zeros((H + 2, W + 2), dtype=bool)
    fg[1:-1, 1:-1] = (np.asarray(mask) != 0)
    ys, xs = np.nonzero(fg)
    # ---- (i) states (p, s): neighbour in direction s is foreground ("arrived at p from there")
    sid = -np.ones((H + 2, W + 2, 8), dtype=np.int64)
    st_y, st_x, st_s = [], [], []
    for s in range(8):
        ok = fg[ys + DY[s], xs + DX[s]]
        st_y.append(ys[ok]); st_x.append(xs[ok]); st_s.append(np.full(ok.sum(), s))
    st_y = np.concatenate(st_y); st_x = np.concatenate(st_x); st_s = np.concatenate(st_s)
    n = len(st_y)
    sid[st_y, st_x, st_s] = np.arange(n)
    nxt = np.empty(n, dtype=np.int64); neg = np.zeros(n, dtype=bool); emit = np.zeros(n, dtype=bool)
    undecided = np.ones(n, dtype=bool)
    for t in range(1, 9):
        s2 = (st_s + t) & 7
        hit = undecided & fg[st_y + DY_arr[s2], st_x + DX_arr[s2]]
        ny, nx = st_y[hit] + DY_arr[s2[hit]], st_x[hit] + DX_arr[s2[hit]]
        nxt[hit] = sid[ny, nx, (s2[hit] + 4) & 7]
        neg[hit] = (st_s[hit] != 0) & (st_s[hit] + t >= 9)
        emit[hit] = s2[hit] != ((st_s[hit] + 4) & 7)
        undecided &= ~hit
    assert not undecided.any()
    # ---- cycles by pointer doubling (label = smallest state id of the cycle)
    lab = np.arange(n); j = nxt.copy()
    for _ in range(max(1, int(np.ceil(np.log2(max(n, 2)))) + 1)):
        lab = np.minimum(lab, lab[j]); j = j[j]
    # per pixel: cycles through it and whether the cycle marks it negative
    marks = {}
    for k in range(n):
        d = marks.setdefault((st_y[k], st_x[k]), {})
        d[lab[k]] = d.get(lab[k], False) or bool(neg[k])
    # ---- candidates in raster order and their start states
    def start_state(y, x, s0):
        s = s0
        for _ in range(8):
            s = (s - 1) & 7
            if fg[y + DY[s], x + DX[s]]:
                return sid[y, x, s]
        return -1
    order = np.lexsort((xs, ys))
    started = {}        # cycle label (or ('iso', y, x)) -> dict(nbd, is_hole, parent, start)
    nbd_of = {}         # nbd -> record
    nbd = 1
    nbd_of[1] = dict(is_hole=True, parent=0, cyc=None)
    def value(y, x):
        """f(y, x) right now from the started cycles through the pixel."""
        d = marks.get((y, x), {})
        best_neg, first_pos = None, None
        for c, isneg in d.items():
            if c in started:
                b = started[c]["nbd"]
                if isneg:
                    best_neg = b if best_neg is None else max(best_neg, b)
                first_pos = b if first_pos is None else min(first_pos, b)
        if ('iso', y, x) in started:
            return -started[('iso', y, x)]["nbd"]
        if best_neg is not None: return -best_neg
        if first_pos is not None: return first_pos
        return 1
    # sequential over pixels of each row that lie on a border or are candidates (event level)
    border_px = set(marks.keys())
    cur_row, lnbd = -1, 1
    for k in order:
        y, x = int(ys[k]), int(xs[k])
        if y != cur_row:
            cur_row, lnbd = y, 1
        west0, east0 = not fg[y, x - 1], not fg[y, x + 1]
        if (y, x) not in border_px and not west0 and not east0:
            continue
        f = value(y, x)
        is_hole = None
        if f == 1 and west0:
            is_hole, s0 = False, 4
        elif f >= 1 and east0:
            is_hole, s0 = True, 0
            if f > 1: lnbd = f
        if is_hole is not None:
            ss = start_state(y, x, s0)
            key = ('iso', y, x) if ss < 0 else lab[ss]
            assert key not in started, "a border would be traced twice"
            nbd += 1
            b0 = nbd_of[lnbd]
            parent = (lnbd if not b0["is_hole"] else b0["parent"]) if is_hole else (lnbd if b0["is_hole"] else b0["parent"])
            rec = dict(nbd=nbd, is_hole=is_hole, parent=parent, start=ss, px=(x - 1, y - 1))
            started[key] = rec; nbd_of[nbd] = rec
        f = value(y, x)
        if f != 1: lnbd = abs(f)
    # ---- points: walk each started cycle from its start state, keep the emitting states
    out_b = sorted(nbd_of)[1:]
    children = {}
    for b in out_b:
        children.setdefault(nbd_of[b]["parent"], []).insert(0, b)
    seq = []
    def walk(p):
        for c in children.get(p, []):
            seq.append(c); walk(c)
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    walk(1)
    contours = []
    for b in seq:
        r = nbd_of[b]
        if r["start"] < 0:
            contours.append(np.array([r["px"]], dtype=np.int32).reshape(-1, 1, 2)); continue
        pts, k = [], r["start"]
        while True:
            if (not simple) or emit[k]: pts.append((st_x[k] - 1, st_y[k] - 1))
            k = nxt[k]
            if k == r["start"]: break
        contours.append(np.array(pts, dtype=np.int32).reshape(-1, 1, 2))
    idx = {b: i for i, b in enumerate(seq)}
    hier = np.full((len(seq), 4), -1, dtype=np.int32)
    for p, ch in children.items():
        for i, c in enumerate(ch):
            if i + 1 < len(ch): hier[idx[c], 0] = idx[ch[i + 1]]
            if i > 0: hier[idx[c], 1] = idx[ch[i - 1]]
            hier[idx[c], 3] = idx[p] if p in idx else -1
        if p in idx and ch: hier[idx[p], 2] = idx[ch[0]]
    return contours, hier



# ---------------------------------------------------------------------------------------------
# The same with step (iii) — the only sequential part above — as a FIXED POINT, which is the form
# pychips_b200/csrc/contours_core.h runs on the device (StartRound / StartCommit):
#   start[c] = the first candidate of cycle c, in raster order, that is eligible given the starts
#   of the OTHER cycles through that pixel as of the previous round (all unset at first);
#   outer candidate (west 0): eligible iff no cycle through the pixel has started earlier;
#   hole candidate (east 0): eligible iff the outer rule does not fire there and no earlier-started
#   cycle marks the pixel negative.
# A start depends only on starts at earlier raster positions, so the raster scan's answer is the
# unique fixed point; Jacobi rounds reach it in 2-7 rounds on noise and O(image side) rounds on a
# lattice of 1-pixel walls (`stats` receives the number of rounds).  Border numbers are then the
# ranks of the start positions and the parent of a border follows from the first marked pixel to
# its left in the row, evaluated "as of" that pixel's own scan time.

def find_contours_fixed_point(mask, simple=True, stats=None):
    H, W = mask.shape
    fg = np.zeros((H + 2, W + 2), dtype=bool)
    fg[1:-1, 1:-1] = (np.asarray(mask) != 0)
    ys, xs = np.nonzero(fg)
    sid = -np.ones((H + 2, W + 2, 8), dtype=np.int64)
    st_y, st_x, st_s = [], [], []
    for s in range(8):
        ok = fg[ys + DY[s], xs + DX[s]]
        st_y.append(ys[ok]); st_x.append(xs[ok]); st_s.append(np.full(ok.sum(), s))
    st_y = np.concatenate(st_y); st_x = np.concatenate(st_x); st_s = np.concatenate(st_s)
    n = len(st_y)
    sid[st_y, st_x, st_s] = np.arange(n)
    nxt = np.empty(n, dtype=np.int64); neg = np.zeros(n, dtype=bool); emit = np.zeros(n, dtype=bool)
    und = np.ones(n, dtype=bool)
    for t in range(1, 9):
        s2 = (st_s + t) & 7
        hit = und & fg[st_y + DY_arr[s2], st_x + DX_arr[s2]]
        nxt[hit] = sid[st_y[hit] + DY_arr[s2[hit]], st_x[hit] + DX_arr[s2[hit]], (s2[hit] + 4) & 7]
        neg[hit] = (st_s[hit] != 0) & (st_s[hit] + t >= 9)
        emit[hit] = s2[hit] != ((st_s[hit] + 4) & 7)
        und &= ~hit
    lab = np.arange(n); j = nxt.copy()
    for _ in range(max(1, int(np.ceil(np.log2(max(n, 2)))) + 1)):
        lab = np.minimum(lab, lab[j]); j = j[j]
    pos = lambda y, x: y * (W + 2) + x
    # per pixel: {cycle: neg}
    marks = {}
    for k in range(n):
        d = marks.setdefault((int(st_y[k]), int(st_x[k])), {})
        d[int(lab[k])] = d.get(int(lab[k]), False) or bool(neg[k])
    def start_state(y, x, s0):
        s = s0
        for _ in range(8):
            s = (s - 1) & 7
            if fg[y + DY[s], x + DX[s]]:
                return int(sid[y, x, s])
        return -1
    # candidates per cycle (raster order); isolated pixels are their own cycles ('iso')
    cand = {}          # cycle -> list of (pos, y, x, is_hole, start_state)
    for k in np.lexsort((xs, ys)):
        y, x = int(ys[k]), int(xs[k])
        w0, e0 = not fg[y, x - 1], not fg[y, x + 1]
        if w0:
            ss = start_state(y, x, 4)
            key = ('iso', y, x) if ss < 0 else int(lab[ss])
            cand.setdefault(key, []).append((pos(y, x), y, x, False, ss))
        if e0:
            ss = start_state(y, x, 0)
            key = ('iso', y, x) if ss < 0 else int(lab[ss])
            cand.setdefault(key, []).append((pos(y, x), y, x, True, ss))
    def cycles_at(y, x):
        d = dict(marks.get((y, x), {}))
        if ('iso', y, x) in cand: d[('iso', y, x)] = True       # an isolated pixel is marked negative
        return d
    # ---- fixed point: start[c] = (pos, is_hole, start_state) of the first eligible candidate
    start = {c: None for c in cand}
    def started_before(c, p, strict=True):
        s = start.get(c)
        return s is not None and (s[0] < p if strict else s[0] <= p)
    def eligible(c, p, y, x, is_hole):
        here = cycles_at(y, x)
        earlier = [cc for cc in here if started_before(cc, p)]
        w0 = not fg[y, x - 1]
        if not is_hole:
            return len(earlier) == 0
        if w0 and len(earlier) == 0:
            return False                      # the outer branch fires at this pixel instead
        return not any(here[cc] for cc in earlier)
    rounds = 0
    while True:
        rounds += 1
        new = {}
        for c, lst in cand.items():
            new[c] = None
            for (p, y, x, ih, ss) in lst:
                if eligible(c, p, y, x, ih):
                    new[c] = (p, ih, ss, y, x)
                    break
        if new == start: break
        start = new
        assert rounds < 200
    if stats is not None: stats.append(rounds)
    real = [(v[0], c) for c, v in start.items() if v is not None]
    real.sort()
    nbd_of_cycle = {c: i + 2 for i, (_, c) in enumerate(real)}
    info = {1: dict(is_hole=True, parent=0)}
    # ---- parents: value of a pixel at a given time
    def value_at(y, x, time_pos):                 # cycles through (y,x) started at positions <= time_pos
        here = cycles_at(y, x)
        st = [(start[c][0], c) for c in here if start.get(c) is not None and start[c][0] <= time_pos]
        if not st: return 1
        negs = [b for b in st if here[b[1]]]
        if negs: return -nbd_of_cycle[max(negs)[1]]
        return nbd_of_cycle[min(st)[1]]
    row_border = {}
    for (y, x) in set(list(marks.keys()) + [(k[1], k[2]) for k in cand if isinstance(k, tuple)]):
        row_border.setdefault(y, []).append(x)
    for y in row_border: row_border[y].sort()
    for p, c in real:
        _, ih, ss, y, x = start[c]
        lnbd = None
        if ih:
            f = value_at(y, x, p - 1)
            if f > 1: lnbd = f
        if lnbd is None:
            lnbd = 1
            for qx in reversed([q for q in row_border.get(y, []) if q < x]):
                f = value_at(y, qx, pos(y, qx))
                if f != 1:
                    lnbd = abs(f); break
        b0 = info[lnbd]
        parent = (lnbd if not b0["is_hole"] else b0["parent"]) if ih else (lnbd if b0["is_hole"] else b0["parent"])
        info[nbd_of_cycle[c]] = dict(is_hole=ih, parent=parent, start=ss, px=(x - 1, y - 1))
    seq_b = sorted(info)[1:]
    children = {}
    for b in seq_b: children.setdefault(info[b]["parent"], []).insert(0, b)
    seq = []
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    def walk(pn):
        for cc in children.get(pn, []):
            seq.append(cc); walk(cc)
    walk(1)
    contours = []
    for b in seq:
        r = info[b]
        if r["start"] < 0:
            contours.append(np.array([r["px"]], dtype=np.int32).reshape(-1, 1, 2)); continue
        pts, k = [], r["start"]
        while True:
            if (not simple) or emit[k]: pts.append((st_x[k] - 1, st_y[k] - 1))
            k = nxt[k]
            if k == r["start"]: break
        contours.append(np.array(pts, dtype=np.int32).reshape(-1, 1, 2))
    idx = {b: i for i, b in enumerate(seq)}
    hier = np.full((len(seq), 4), -1, dtype=np.int32)
    for pn, ch in children.items():
        for i, cc in enumerate(ch):
            if i + 1 < len(ch): hier[idx[cc], 0] = idx[ch[i + 1]]
            if i > 0: hier[idx[cc], 1] = idx[ch[i - 1]]
            hier[idx[cc], 3] = idx[pn] if pn in idx else -1
        if pn in idx and ch: hier[idx[pn], 2] = idx[ch[0]]
    return contours, hier
