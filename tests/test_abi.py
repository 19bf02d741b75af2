"""CPU: the C-ABI library loads, exports every symbol `include/chips_b200.h` declares, and the
ctypes mirrors of its structs match the C layout (checked with gcc).  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HDR = os.path.join(ROOT, "include", "chips_b200.h")


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long|size_t)\s+(chips_\w+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from pychips_b200 import _lib, build
    build.build_library()
    lib = C.CDLL(build.LIB)
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in chips_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(names), set(_lib.SIGNATURES) ^ set(names)
    assert _lib.load().chips_version() == 100


def test_struct_layouts_match_c(tmp_path):
    from pychips_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "chips_b200.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(chips_result), sizeof(chips_plan),'
        ' offsetof(chips_result, n_samples), offsetof(chips_result, median_errors),'
        ' offsetof(chips_plan, off_order), offsetof(chips_plan, bytes));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [C.sizeof(_lib.Result), C.sizeof(_lib.Plan), _lib.Result.n_samples.offset,
           _lib.Result.median_errors.offset, _lib.Plan.off_order.offset, _lib.Plan.bytes.offset]
    assert [int(v) for v in out] == got


def test_plan_init_geometry_and_argument_errors():
    from pychips_b200 import _lib
    lib = _lib.load()
    p = _lib.Plan()
    assert lib.chips_plan_init(C.byref(p), 4096, 4096, 51, 500, 20, 4, 2048, 2048, 1577) == 0
    assert (p.roi_x0, p.roi_y0, p.roi_w, p.roi_h) == (471, 471, 3155, 3155)
    assert p.masked == 1 and p.bytes > 0 and p.bp_pitch % 16 == 0
    offs = [getattr(p, f) for f, _ in _lib.Plan._fields_ if f.startswith("off_")]
    assert all(o % 256 == 0 for o in offs) and offs == sorted(offs)
    # unmasked (synoptic): whole image, no union-find buffers
    assert lib.chips_plan_init(C.byref(p), 1440, 3600, 3, 5000, 25, 4, 0, 0, -1) == 0
    assert (p.roi_w, p.roi_h, p.masked) == (3600, 1440, 0)
    assert p.off_parent_b == p.off_area == p.off_ovf < p.bytes
    assert p.blk_rows % 16 == 0 and p.blk_nbx == 29 and p.blk_nby * p.blk_rows >= 1440
    # disk clipped by the image border
    assert lib.chips_plan_init(C.byref(p), 100, 100, 5, 10, 3, 8, 50, 50, 80) == 0
    assert (p.roi_x0, p.roi_y0, p.roi_w, p.roi_h) == (0, 0, 100, 100)
    for bad in [(64, 64, 4, 10, 3, 4), (64, 64, 77, 10, 3, 4), (64, 64, 5, 0, 3, 4),
                (64, 64, 5, 10, 0, 4), (64, 64, 5, 10, 65, 4), (64, 64, 5, 10, 3, 2), (0, 64, 5, 10, 3, 4)]:
        assert lib.chips_plan_init(C.byref(p), *bad, 32, 32, 20) == -1, bad
    assert lib.chips_plan_init(None, 64, 64, 5, 10, 3, 4, 32, 32, 20) == -1


def test_null_pointers_are_rejected_without_touching_the_gpu():
    from pychips_b200 import _lib
    lib = _lib.load()
    p = _lib.Plan()
    lib.chips_plan_init(C.byref(p), 64, 64, 5, 10, 3, 4, 32, 32, 20)
    assert lib.chips_medfilt2d(None, None, C.byref(p), 32, 32, 20, None, None) == -1
    assert lib.chips_histogram(None, C.byref(p), 32, 32, 20, None, None) == -1
    assert lib.chips_threshold_prob(None, C.byref(p), 32, 32, 20, 0.8, None, None) == -1
    assert lib.chips_cleanup(C.byref(p), 32, 32, 20, 1.0, 1e-3, None, None) == -1
    offs = (C.c_double * 3)(0.0, 1.0, 2.0)
    assert lib.chips_detect(None, C.byref(p), 32, 32, 20, 5.0, float("nan"), offs, 0.8, 1e-3,
                            None, None, None) == -1
    assert lib.chips_select_threshold(C.byref(p), 5.0, float("nan"), None, None, None) == -1
    assert lib.chips_select_threshold(C.byref(p), 5.0, float("nan"), offs, None, None) == -1
    assert lib.chips_medfilt2d_bruteforce(None, None, 8, 8, 3, 4, 0, 0, -1, None) == -1


@pytest.mark.parametrize("H,W,k,r", [(4096, 4096, 51, 1577), (1024, 1024, 11, 394), (20, 20, 31, -1),
                                     (1, 1, 3, -1), (9, 300, 11, -1), (1440, 3600, 3, -1), (300, 9, 73, -1),
                                     (512, 512, 71, 300)])
def test_plan_block_geometry_is_consistent(H, W, k, r):
    """The median's 128 x L blocks tile the ROI, L is a multiple of the 16-pixel refine tile, the
    block images hold the halo, and every workspace offset is aligned and inside the blob."""
    from pychips_b200 import _lib
    lib = _lib.load()
    p = _lib.Plan()
    assert lib.chips_plan_init(C.byref(p), H, W, k, 500, 20, 4, W // 2, H // 2, r) == 0
    half = k // 2
    assert p.blk_rows % 16 == 0 and p.blk_rows >= 16
    assert p.blk_nbx * 128 >= p.roi_w and (p.blk_nbx - 1) * 128 < p.roi_w
    assert p.blk_nby * p.blk_rows >= p.roi_h and (p.blk_nby - 1) * p.blk_rows < p.roi_h
    assert p.bp_pitch % 16 == 0 and p.bp_pitch >= 128 + 2 * half + 16 and p.bp_rows == p.blk_rows + 2 * half
    assert 0 <= p.roi_x0 and p.roi_x0 + p.roi_w <= W and 0 <= p.roi_y0 and p.roi_y0 + p.roi_h <= H
    offs = [getattr(p, f) for f, _ in _lib.Plan._fields_ if f.startswith("off_")]
    assert all(o % 256 == 0 and o <= p.bytes for o in offs)
    nblk = p.blk_nbx * p.blk_nby
    assert p.off_bstar - p.off_bucket >= nblk * p.bp_pitch * p.bp_rows
    assert p.off_bucket - p.off_order >= nblk * (128 + 2 * half) * p.bp_rows * 2
    # the block sort holds a halo-extended block in shared memory; 128 buckets of blk_bucket ranks
    nelem = (128 + 2 * half) * p.bp_rows
    assert nelem <= 24576 and p.blk_bucket * 128 >= nelem > p.blk_bucket * 127 - 128


def test_contour_structs_and_argument_errors(tmp_path):
    """chips_contour_rec / chips_contour_counts: the numpy mirrors match the C layout; bad arguments
    are rejected before anything touches the GPU; the workspace grows with the capacity."""
    import numpy as np
    from pychips_b200 import _lib, contours as pc
    prog = tmp_path / "szc.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "chips_b200.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(chips_contour_rec), sizeof(chips_contour_counts),'
        ' offsetof(chips_contour_rec, area2), offsetof(chips_contour_rec, offset),'
        ' offsetof(chips_contour_counts, overflow));return 0;}\n')
    exe = tmp_path / "szc"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out == [pc.REC_DTYPE.itemsize, pc.COUNTS_DTYPE.itemsize, pc.REC_DTYPE.fields["area2"][1],
                   pc.REC_DTYPE.fields["offset"][1], pc.COUNTS_DTYPE.fields["overflow"][1]]
    lib = _lib.load()
    small, big = (lib.chips_contours_workspace_bytes(4096, 4096, c) for c in (1 << 16, 1 << 18))
    assert 0 < small < big and big - small >= (3 << 16) * 8 * 4 * 9
    assert lib.chips_contours_workspace_bytes(0, 10, 100) == 0
    one = C.c_void_p(256)           # never dereferenced: the checks come first
    ok = dict(w=64, h=64, pitch=64, mode=1, thr=0, cap=1000, rounds=8, maxc=10, maxp=10, ws_bytes=1 << 30)

    def call(**kw):
        a = dict(ok, **kw)
        return lib.chips_contours(one, a["w"], a["h"], a["pitch"], a["mode"], a["thr"], a["cap"], a["rounds"],
                                  one, a["maxc"], one, a["maxp"], one, one, a["ws_bytes"], None)
    for bad in (dict(w=0), dict(pitch=63), dict(mode=2), dict(cap=0), dict(cap=1 << 28), dict(rounds=1),
                dict(rounds=(1 << 20) + 1), dict(maxc=0), dict(maxp=0), dict(w=1 << 15, h=1 << 15, pitch=1 << 15)):
        assert call(**bad) == -1, bad
    assert call(ws_bytes=1024) == -3
    assert lib.chips_contours(None, 64, 64, 64, 1, 0, 1000, 8, one, 10, one, 10, one, one, 1 << 30, None) == -1


def test_input_rect_covers_the_median_blocks():
    from pychips_b200 import _lib
    lib = _lib.load()
    p = _lib.Plan()
    rect = (C.c_int32 * 4)()
    assert lib.chips_plan_init(C.byref(p), 4096, 4096, 51, 500, 20, 4, 2048, 2048, 1577) == 0
    assert lib.chips_input_rect(C.byref(p), rect) == 0
    x0, y0, x1, y1 = rect
    assert (x0, y0) == (471 - 25, 471 - 25)
    assert x1 == min(4096, 471 + p.blk_nbx * 128 + 25) and y1 == min(4096, 471 + p.blk_nby * p.blk_rows + 25)
    assert x1 >= 471 + 3155 + 25 and y1 >= 471 + 3155 + 25
    assert (x1 - x0) * (y1 - y0) < 0.66 * 4096 * 4096
    assert lib.chips_plan_init(C.byref(p), 1440, 3600, 3, 5000, 25, 4, 0, 0, -1) == 0     # unmasked: everything
    assert lib.chips_input_rect(C.byref(p), rect) == 0 and tuple(rect) == (0, 0, 3600, 1440)
    assert lib.chips_plan_init(C.byref(p), 100, 100, 5, 10, 3, 8, 50, 50, 80) == 0         # clipped disk
    assert lib.chips_input_rect(C.byref(p), rect) == 0 and tuple(rect) == (0, 0, 100, 100)
    assert lib.chips_input_rect(None, rect) == -1
    # the spans: disjoint ascending row intervals inside the rectangle that cover every on-disk
    # block with its halo (recomputed here from the plan) and, at 4096^2, 52 % of the image
    for (n, k, r) in ((4096, 51, 1577), (1024, 11, 394), (256, 31, 120), (200, 7, 140)):
        c = n // 2
        assert lib.chips_plan_init(C.byref(p), n, n, k, 500, 20, 4, c, c, r) == 0
        assert lib.chips_input_rect(C.byref(p), rect) == 0
        sp = (C.c_int32 * (4 * 4096))()
        ns = lib.chips_input_spans(C.byref(p), c, c, r, sp, 4096)
        assert 1 <= ns <= 200
        spans = [tuple(sp[4 * i:4 * i + 4]) for i in range(ns)]
        import numpy as np
        cover = np.zeros((n, n), bool)
        prev = -1
        for (y0, y1, x0, x1) in spans:
            assert prev <= y0 < y1 and rect[1] <= y0 and y1 <= rect[3] and rect[0] <= x0 < x1 <= rect[2]
            prev = y1
            cover[y0:y1, x0:x1] = True
        half = k // 2
        need = np.zeros((n, n), bool)
        for by in range(p.blk_nby):
            for bx in range(p.blk_nbx):
                X, Y = p.roi_x0 + 128 * bx, p.roi_y0 + p.blk_rows * by
                nx, ny = min(max(c, X), X + 127), min(max(c, Y), Y + p.blk_rows - 1)
                if (nx - c) ** 2 + (ny - c) ** 2 <= r * r:
                    need[max(0, Y - half):Y + p.blk_rows + half, max(0, X - half):X + 128 + half] = True
        assert not (need & ~cover).any()
        if n == 4096:
            assert 0.50 < cover.mean() < 0.53
    assert lib.chips_plan_init(C.byref(p), 1024, 1024, 11, 500, 20, 4, 512, 512, 394) == 0
    assert lib.chips_input_spans(C.byref(p), 512, 512, 394, sp, 1) == -3
    assert lib.chips_plan_init(C.byref(p), 1440, 3600, 3, 5000, 25, 4, 0, 0, -1) == 0
    assert lib.chips_input_spans(C.byref(p), 0, 0, -1, sp, 8) == 1 and tuple(sp[0:4]) == (0, 1440, 0, 3600)
    one = C.c_void_p(256)
    assert lib.chips_copy_rect(one, one, 64, 60, 0, 8, 1, 1, None) == -1      # runs past the pitch
    assert lib.chips_copy_rect(None, one, 64, 0, 0, 8, 1, 1, None) == -1
