"""CPU: oracle/contours_spec.py (Suzuki-Abe border following restated from the publication)
against the installed OpenCV — the pin for the contour lists of SURVEY.md §8(f) row 4."""
import cv2
import numpy as np
import pytest

from oracle import contours_spec as cs
from oracle import chips_oracle as co
from pychips_b200 import synth


def _same(mask, simple):
    ref, hi = cv2.findContours(mask, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE if simple else cv2.CHAIN_APPROX_NONE)
    got, gh, _ = cs.find_contours(mask, simple)
    assert len(ref) == len(got)
    for k, (a, b) in enumerate(zip(ref, got)):
        assert np.array_equal(a, b), k
    if len(ref):
        assert np.array_equal(hi[0], gh)
    return ref


@pytest.mark.parametrize("seed", range(6))
def test_random_noise_masks_points_order_hierarchy(seed):
    rng = np.random.default_rng(seed)
    for _ in range(40):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        m = (rng.random((h, w)) < rng.choice([0.2, 0.5, 0.8])).astype(np.uint8)
        _same(m, True)
        _same(m, False)


def test_more_than_255_contours_and_thin_walls():
    rng = np.random.default_rng(7)
    m = (rng.random((150, 170)) < 0.55).astype(np.uint8)
    assert len(_same(m, True)) > 1000           # border numbers are not wrapped
    ring = np.zeros((9, 9), np.uint8)           # a 1-pixel wall between the outside and a hole
    ring[1:8, 1:8] = 1
    ring[2:7, 2:7] = 0
    ring[4, 4] = 1
    _same(ring, True)
    _same(np.ones((5, 7), np.uint8), True)      # touches every image border
    _same(np.zeros((4, 4), np.uint8), True)


def test_area_and_kept_contours_match_the_reference_step():
    """contourArea (shoelace) and the kept-contour lists / hierarchy rows of
    clean_small_scale_structures on a pipeline mask."""
    img, r = synth.synthetic_disk(256, 3)
    d = synth.FakeDisk(img.astype(np.float64), r)
    co.OracleChips(medfilt_kernel=11).run(d)
    disk_area = np.pi * r ** 2
    for key in list(d.solar_ch_regions.__dict__)[::5]:
        reg = d.solar_ch_regions.__dict__[key]
        raw = reg.raw_map.astype(np.uint8)
        cnts, hier, _ = cs.find_contours(raw, True)
        ref, _ = cv2.findContours(raw, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        for a, b in zip(ref, cnts):
            assert cv2.contourArea(a) == cs.contour_area(b)
        kept, khier, dx = cs.clean_small_scale_structures(cnts, hier, raw.shape, disk_area, 1e-3)
        assert len(kept) == len(reg.contours)
        for a, b in zip(reg.contours, kept):
            assert np.array_equal(a, b)
        assert np.array_equal(np.asarray(reg.hierarchy), khier)
        assert np.array_equal(dx, reg.map)


@pytest.mark.parametrize("seed", range(4))
def test_parallel_formulation_matches_opencv(seed):
    """oracle/contours_parallel_spec.py: cycles of the state permutation + event-level start
    resolution give OpenCV's contours, order and hierarchy."""
    from oracle import contours_parallel_spec as ps
    rng = np.random.default_rng(100 + seed)
    for _ in range(25):
        h, w = int(rng.integers(1, 34)), int(rng.integers(1, 34))
        m = (rng.random((h, w)) < rng.choice([0.3, 0.5, 0.7])).astype(np.uint8)
        for simple in (True, False):
            ref, hi = cv2.findContours(m, cv2.RETR_TREE,
                                       cv2.CHAIN_APPROX_SIMPLE if simple else cv2.CHAIN_APPROX_NONE)
            got, gh = ps.find_contours_parallel(m, simple)
            assert len(ref) == len(got)
            assert all(np.array_equal(a, b) for a, b in zip(ref, got))
            assert len(ref) == 0 or np.array_equal(hi[0], gh)


@pytest.mark.parametrize("seed", range(3))
def test_fixed_point_start_resolution_matches_opencv(seed):
    """oracle/contours_parallel_spec.py::find_contours_fixed_point — the start positions as a
    Jacobi fixed point (the form the device passes use) — against OpenCV, and how many rounds it takes."""
    from oracle import contours_parallel_spec as ps
    rng = np.random.default_rng(200 + seed)
    rounds = []
    for _ in range(25):
        h, w = int(rng.integers(1, 34)), int(rng.integers(1, 34))
        m = (rng.random((h, w)) < rng.choice([0.3, 0.5, 0.7])).astype(np.uint8)
        ref, hi = cv2.findContours(m, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        got, gh = ps.find_contours_fixed_point(m, True, rounds)
        assert len(ref) == len(got)
        assert all(np.array_equal(a, b) for a, b in zip(ref, got))
        assert len(ref) == 0 or np.array_equal(hi[0], gh)
    assert max(rounds) <= 12
    lattice = np.zeros((30, 31), np.uint8)
    lattice[::2, :] = 1
    lattice[:, ::3] = 1
    ref, hi = cv2.findContours(lattice, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    rounds = []
    got, gh = ps.find_contours_fixed_point(lattice, True, rounds)
    assert all(np.array_equal(a, b) for a, b in zip(ref, got)) and np.array_equal(hi[0], gh)
    assert rounds[0] > 8            # O(image side): the device call reports `changed` and is re-enqueued
