"""CPU: the passes of pychips_b200/csrc/contours_core.h (the code contours.cu launches as CUDA
kernels), run as host loops by tests/emu/contours_emu.cpp, against the installed OpenCV — contours,
order, hierarchy and areas of cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE).  The GPU parity
test of the same entry point is tests/test_gpu_parity.py::test_contours_*."""
import ctypes
import os
import subprocess

import cv2
import numpy as np
import pytest

from pychips_b200 import contours as pc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "contours_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                    os.path.join(HERE, "emu", "contours_emu.cpp")], check=True)
    lib = ctypes.CDLL(out)
    lib.emu_contours_workspace_bytes.restype = ctypes.c_size_t
    lib.emu_contours_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong]
    lib.emu_set_order.argtypes = [ctypes.c_int]
    lib.emu_contours.restype = ctypes.c_int
    lib.emu_contours.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_size_t]
    return lib


def run_emu(lib, img, mode=1, thr=0, cap=None, rounds=12, max_contours=None, max_points=None):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    cap = cap or max(1, h * w)
    max_contours = max_contours or cap
    max_points = max_points or 8 * cap
    nbytes = lib.emu_contours_workspace_bytes(w, h, cap)
    ws = np.full(nbytes, 0xA5, dtype=np.uint8)               # garbage: the passes must initialise
    recs = np.zeros(max_contours, dtype=pc.REC_DTYPE)
    points = np.full(2 * max_points, -7, dtype=np.int32)
    counts = np.zeros(1, dtype=pc.COUNTS_DTYPE)
    rc = lib.emu_contours(img.ctypes.data, w, h, w, mode, thr, cap, rounds, recs.ctypes.data, max_contours,
                          points.ctypes.data, max_points, counts.ctypes.data, ws.ctypes.data, nbytes)
    assert rc == 0
    return recs, points, counts[0]


def check(lib, mask, **kw):
    ref, hi = cv2.findContours(mask, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    recs, points, c = run_emu(lib, mask, **kw)
    assert c["overflow"] == 0 and c["changed"] == 0
    assert c["n_contours"] == len(ref)
    got, hier, area2 = pc.assemble(recs[:c["n_contours"]], points[:2 * c["n_points"]])
    for k, (a, b) in enumerate(zip(ref, got)):
        assert np.array_equal(a, b), k
    if len(ref):
        assert np.array_equal(hi[0], hier)
        assert np.array_equal(area2, [int(round(2 * cv2.contourArea(a))) for a in ref])
    return ref


@pytest.mark.parametrize("seed", range(6))
def test_random_masks(emu, seed):
    rng = np.random.default_rng(seed)
    for _ in range(40):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        m = (rng.random((h, w)) < rng.choice([0.2, 0.5, 0.8])).astype(np.uint8)
        check(emu, m)


def test_structured_masks(emu):
    full = np.ones((20, 33), np.uint8)
    check(emu, full)
    check(emu, np.zeros((5, 9), np.uint8))
    ring = np.zeros((9, 9), np.uint8)
    ring[1:8, 1:8] = 1
    ring[2:7, 2:7] = 0
    ring[4, 4] = 1
    check(emu, ring)
    yy, xx = np.mgrid[:120, :130]
    blobs = ((np.sin(xx / 7.0) * np.cos(yy / 5.0) > 0.2) | ((xx - 60) ** 2 + (yy - 60) ** 2 < 400)).astype(np.uint8)
    blobs[50:70, 55:65] = 0
    blobs[58:62, 58:61] = 1
    assert len(check(emu, blobs)) > 10
    rng = np.random.default_rng(5)
    big = (rng.random((150, 170)) < 0.55).astype(np.uint8)
    assert len(check(emu, big)) > 1000


def test_threshold_mode_and_pitch(emu):
    rng = np.random.default_rng(9)
    lev = rng.integers(0, 6, size=(31, 45)).astype(np.uint8)
    for t in range(6):
        mask = (lev <= t).astype(np.uint8)
        ref, hi = cv2.findContours(mask, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        recs, points, c = run_emu(emu, lev, mode=0, thr=t)
        got, hier, _ = pc.assemble(recs[:c["n_contours"]], points[:2 * c["n_points"]])
        assert len(got) == len(ref) and all(np.array_equal(a, b) for a, b in zip(ref, got))
        assert np.array_equal(hi[0], hier)


def test_overflow_is_reported(emu):
    rng = np.random.default_rng(2)
    m = (rng.random((40, 40)) < 0.5).astype(np.uint8)
    _, _, c = run_emu(emu, m, cap=16)
    assert c["overflow"] & 1 and c["n_contours"] == 0 and c["n_border_px"] > 16
    _, _, c = run_emu(emu, m, max_contours=3)
    assert c["overflow"] & 2
    _, _, c = run_emu(emu, m, max_points=5)
    assert c["overflow"] & 4
    _, _, c = run_emu(emu, m, rounds=2)
    assert c["overflow"] == 0      # `changed` tells whether two rounds were enough


def test_output_order_host_logic():
    # frame -> 2 (outer) -> 3 (hole) -> 4 (outer) ; 5 (outer, top level)
    seq, hier = pc.output_order(np.array([1, 2, 3, 1]))
    assert seq.tolist() == [3, 0, 1, 2]
    assert hier.tolist() == [[1, -1, -1, -1], [-1, 0, 2, -1], [-1, -1, 3, 1], [-1, -1, -1, 2]]


def test_lattice_of_one_pixel_walls_needs_many_start_rounds(emu):
    """The start positions of a lattice settle in O(image side) Jacobi rounds: `changed` says when
    the enqueued rounds were too few, and any number of rounds is allowed (rotating flag slots)."""
    n = 80
    m = np.zeros((n, n), np.uint8)
    m[::2, :] = 1
    m[:, ::3] = 1
    _, _, c = run_emu(emu, m, rounds=8)
    assert c["changed"] == 1 and c["overflow"] == 0
    check(emu, m, rounds=64)
    check(emu, m, rounds=301)


def test_pipeline_masks_every_threshold(emu):
    """The raw threshold masks of a synthetic detection (what chips.py:382 hands to findContours),
    through the level image + threshold form of the call (mode 0), and the area cut of
    clean_small_scale_structures on the traced lists."""
    from oracle import chips_oracle as co
    from pychips_b200 import synth
    n = 256
    img, r = synth.synthetic_disk(n, 5)
    d = synth.FakeDisk(img.astype(np.float64), r)
    co.OracleChips(medfilt_kernel=11).run(d, keep_contours=True)
    regions = list(vars(d.solar_ch_regions).values())
    filt = d.solar_filter.filt_disk * d.solar_mask.n_mask          # NaN off-disk
    level = np.full((n, n), len(regions), dtype=np.uint8)
    for t in reversed(range(len(regions))):
        level[np.nan_to_num(filt, nan=np.inf) <= regions[t].lim] = t
    disk_area = np.pi * r ** 2
    for t, reg in enumerate(regions):
        recs, points, c = run_emu(emu, level, mode=0, thr=t)
        assert c["overflow"] == 0 and c["changed"] == 0
        got, hier, area2 = pc.assemble(recs[:c["n_contours"]], points[:2 * c["n_points"]])
        keep = np.flatnonzero(area2 / 2.0 / disk_area >= 1e-3)
        assert len(keep) == len(reg.contours), t
        for i, want in zip(keep, reg.contours):
            assert np.array_equal(got[i], want), t
        assert np.array_equal(np.array([hier[i] for i in keep]), np.asarray(reg.hierarchy)), t


@pytest.mark.parametrize("seed", range(4))
def test_drawn_shapes_nested_and_touching(emu, seed):
    """Rectangles, diagonal strokes, rings and punched holes drawn over each other (nesting several
    levels deep, borders sharing pixels, shapes touching the image frame)."""
    rng = np.random.default_rng(900 + seed)
    for _ in range(12):
        h, w = int(rng.integers(8, 72)), int(rng.integers(8, 72))
        m = np.zeros((h, w), np.uint8)
        for _ in range(int(rng.integers(3, 14))):
            kind = rng.integers(0, 4)
            y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
            y1, x1 = int(rng.integers(y0, h)) + 1, int(rng.integers(x0, w)) + 1
            val = int(rng.integers(0, 2))
            if kind == 0:
                m[y0:y1, x0:x1] = val
            elif kind == 1:                                  # ring, one pixel wide
                m[y0:y1, x0:x1] = 1
                m[y0 + 1:y1 - 1, x0 + 1:x1 - 1] = 0
            elif kind == 2:                                  # diagonal stroke
                n = min(y1 - y0, x1 - x0)
                idx = np.arange(n)
                m[y0 + idx, (x0 + idx) if rng.random() < 0.5 else (x1 - 1 - idx)] = val
            else:                                            # salt
                pts = rng.random((y1 - y0, x1 - x0)) < 0.15
                m[y0:y1, x0:x1][pts] = val
        check(emu, m, rounds=64)


def test_output_order_levelwise_equals_walk():
    """contours.output_order: the level-by-level pre-order (numpy passes) equals the plain tree walk
    on random forests, including chains deeper than its depth limit (fallback)."""
    rng = np.random.default_rng(4)
    for n in (1, 2, 7, 50, 400, 3000):
        for bias in (0.0, 0.5, 0.95):
            parent = np.full(n, 1, dtype=np.int64)                  # border numbers; 1 = frame
            for b in range(1, n):
                if rng.random() < bias:
                    parent[b] = b + 1                               # child of the previous border: deep chains
                elif rng.random() < 0.7:
                    parent[b] = int(rng.integers(0, b)) + 2
            seq, hier = pc.output_order(parent)
            p0 = parent - 2
            b = np.arange(n)
            by_parent = np.lexsort((-b, p0))
            first_child = np.full(n + 1, -1, dtype=np.int64)
            next_sib = np.full(n, -1, dtype=np.int64)
            ps = p0[by_parent]
            head = np.ones(n, dtype=bool)
            head[1:] = ps[1:] != ps[:-1]
            first_child[np.where(ps[head] < 0, n, ps[head])] = by_parent[head]
            same = ~head[1:]
            next_sib[by_parent[:-1][same]] = by_parent[1:][same]
            want = pc._preorder_walk(first_child, next_sib, n)
            assert np.array_equal(seq, want), (n, bias)
            pos = np.empty(n, dtype=np.int64)
            pos[seq] = np.arange(n)
            assert np.array_equal(hier[:, 3], np.where(p0[seq] >= 0, pos[np.maximum(p0[seq], 0)], -1))
    with pytest.raises(RuntimeError):
        pc.output_order(np.array([3, 2]))                           # two borders that are each other's parent


def test_passes_do_not_depend_on_the_order_of_their_indices(emu):
    """Every pass visited in descending and in a pseudo-random index order gives byte-identical
    records and points: no pass reads what another index of the same pass writes (which would be
    a race between CUDA threads)."""
    rng = np.random.default_rng(11)
    masks = [(rng.random((60, 70)) < p).astype(np.uint8) for p in (0.3, 0.55, 0.8)]
    lattice = np.zeros((40, 41), np.uint8)
    lattice[::2, :] = 1
    lattice[:, ::3] = 1
    masks.append(lattice)
    try:
        for m in masks:
            outs = []
            for order in (0, 1, 2):
                emu.emu_set_order(order)
                recs, points, c = run_emu(emu, m, rounds=64)
                n, k = int(c["n_contours"]), int(c["n_points"])
                assert c["changed"] == 0 and c["overflow"] == 0
                outs.append((recs[:n].tobytes(), points[:2 * k].tobytes(), n, k))
            assert outs[0] == outs[1] == outs[2]
    finally:
        emu.emu_set_order(0)
    check(emu, masks[1])
