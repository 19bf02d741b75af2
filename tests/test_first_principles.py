"""CPU: the first-principles restatements (oracle/first_principles.py — the executable spec of
the CUDA kernels) against the third-party calls the reference makes (scipy / numpy / OpenCV)
and against the golden outputs of the reference."""
import os
import re

import cv2
import numpy as np
import pytest
from scipy import ndimage, signal

from conftest import ROOT, load_golden
from oracle import chips_oracle as co
from oracle import first_principles as fp


# ----------------------------------------------------------------------------- circle
@pytest.mark.parametrize("n,r", [(64, 20), (200, 77), (512, 197), (1024, 394), (100, 60), (33, 1)])
def test_circle_is_integer_test(n, r):
    canvas = np.zeros((n, n), dtype=np.float64)
    cv2.circle(canvas, (int(n / 2), int(n / 2)), r, 255, -1)
    assert np.array_equal(canvas > 0, fp.circle_mask(n, n, int(n / 2), int(n / 2), r))


# ----------------------------------------------------------------------------- median
@pytest.mark.parametrize("k", [1, 3, 5, 11, 31])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_median_is_zero_padded_rank(k, dtype):
    rng = np.random.default_rng(k)
    img = rng.gamma(2.0, 30.0, size=(47, 61)).astype(dtype)
    img[5:9, 10:30] = 0.0
    img[20:25, :] = -3.5
    assert np.array_equal(signal.medfilt2d(img, k), fp.median_zero_pad(img, k))


# ----------------------------------------------------------------------------- histogram
@pytest.mark.parametrize("nbins", [7, 500, 5000])
def test_histogram_uniform_matches_numpy(nbins):
    rng = np.random.default_rng(nbins)
    for x in (rng.gamma(2.0, 30.0, 20000), rng.normal(100, 5, 7777).astype(np.float32).astype(np.float64),
              np.linspace(0, 1, 1001), np.array([3.0, 3.0, 3.0]), np.arange(64, dtype=np.float64)):
        c, e = fp.histogram_uniform(x, nbins)
        cn, en = np.histogram(x, bins=nbins)
        assert np.array_equal(c, cn) and np.array_equal(e, en)
        hn, _ = np.histogram(x, bins=nbins, density=True)
        assert np.array_equal(fp.density_from_counts(c, e), hn)


def test_find_peaks_matches_scipy():
    rng = np.random.default_rng(5)
    for trial in range(200):
        n = int(rng.integers(3, 60))
        x = rng.integers(0, 5, n).astype(np.float64)     # many plateaus
        h = float(rng.integers(0, 4))
        ref, _ = signal.find_peaks(x, height=h)
        assert np.array_equal(ref, fp.find_peaks_height(x, h)), (x, h)


def test_limits_rounding():
    rng = np.random.default_rng(1)
    for l0 in rng.uniform(-5, 200, 50):
        for tr in ([0, 20], [-3, 25], [-5, 20]):
            assert np.array_equal(np.round(l0 + np.arange(*tr), 4), fp.limits_from(l0, tr))


# ----------------------------------------------------------------------------- probability
def test_prob_table_matches_calculate_prob():
    rng = np.random.default_rng(9)
    v = np.concatenate([rng.normal(28, 3, 30000), rng.normal(130, 15, 50000)])
    v = v.astype(np.float32).astype(np.float64)
    limit_0 = 28.4339
    lims = fp.limits_from(limit_0, [-3, 25])
    tab = fp.prob_counts(v, limit_0, lims)
    got = fp.prob_from_counts(tab)
    for t, lim in enumerate(lims):
        assert co.calculate_prob(v.copy(), [lim, limit_0]) == got[t]


def test_prob_bins_from_d_cuts():
    rng = np.random.default_rng(3)
    d = np.concatenate([rng.uniform(-30, 30, 400000), rng.normal(0, 0.5, 400000)])
    ps = 1.0 / (1.0 + np.exp(d).round(4))
    h_ref, _ = np.histogram(ps, bins=np.linspace(0, 1, 21))
    cuts = fp.prob_d_cuts()
    b = (d[:, None] <= cuts[None, 1:]).sum(1)
    assert np.array_equal(np.bincount(b, minlength=20), h_ref)


def test_prob_cut_constants_in_cuda_source():
    src = open(os.path.join(ROOT, "pychips_b200", "csrc", "stats.cu")).read()
    body = src[src.index("kProbCut[20] = {") + len("kProbCut[20] = {"):]
    body = body[:body.index("};")]
    vals = [float.fromhex(v) if "x" in v else float(v)
            for v in re.findall(r"-?0x[0-9a-f.]+p[-+]?\d+|0\.0", body)]
    assert len(vals) == 20
    assert np.array_equal(np.array(vals)[1:], fp.prob_d_cuts()[1:])


def test_pairwise_sum_matches_numpy():
    rng = np.random.default_rng(4)
    for n in (1, 4, 7, 8, 9, 16, 20, 23):
        a = rng.uniform(0, 1e7, n)
        assert fp._pairwise_sum(a) == np.sum(a)


# ----------------------------------------------------------------------------- cleanup
def _reference_cleanup(mask_u8, pixel_radius, area_threshold):
    contours, hierarchy = cv2.findContours(mask_u8, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    if len(contours) == 0:
        return np.zeros(mask_u8.shape, np.uint8)
    _, _, dx = co.clean_small_scale_structures(contours, hierarchy, np.zeros(mask_u8.shape),
                                               np.pi * pixel_radius ** 2, area_threshold)
    return dx.astype(np.uint8)


def _random_nested_levels(rng, n, T):
    """Level image with nested masks, holes, islands, border-touching blobs, specks."""
    f = ndimage.gaussian_filter(rng.normal(size=(n, n)), rng.uniform(1.0, 4.0))
    f = (f - f.min()) / (f.max() - f.min())
    lev = np.clip((f * (T + 6)).astype(np.int64) - 3, 0, T)
    speck = rng.random((n, n)) < 0.02
    lev[speck] = rng.integers(0, T + 1, speck.sum())
    return lev.astype(np.uint8)


@pytest.mark.parametrize("seed", range(12))
def test_level_cleanup_equals_findcontours_pipeline(seed):
    rng = np.random.default_rng(seed)
    n, T = int(rng.integers(24, 72)), int(rng.integers(1, 9))
    lev = _random_nested_levels(rng, n, T)
    r = n // 3
    thr = float(rng.choice([0.0, 1e-3, 5e-3, 2e-2]))
    W, maps, labels, areas = fp.cleanup_levels(lev, T, r, thr)
    keep_prev = None
    for t in range(T):
        ref = _reference_cleanup((lev <= t).astype(np.uint8), r, thr)
        assert np.array_equal(maps[t], ref), (seed, t)
        if keep_prev is not None:                      # final maps are nested in t
            assert np.all(maps[t] >= keep_prev)
        keep_prev = maps[t]
        # per-label 2*area equals 2*cv2.contourArea of the top-level contours
        contours, hier = cv2.findContours((lev <= t).astype(np.uint8), cv2.RETR_TREE,
                                          cv2.CHAIN_APPROX_SIMPLE)
        top = sorted(int(round(2 * cv2.contourArea(c))) for i, c in enumerate(contours)
                     if hier[0][i][3] < 0)
        assert top == sorted(int(a) for a in areas[t][1:])


def test_fill_levels_is_holefill_of_each_mask():
    rng = np.random.default_rng(77)
    lev = _random_nested_levels(rng, 64, 6)
    W = fp.fill_levels(lev, 6)
    for t in range(6):
        assert np.array_equal(W <= t, ndimage.binary_fill_holes(lev <= t))


# ----------------------------------------------------------------------------- goldens
@pytest.mark.parametrize("name", ["disk_n256_k11", "disk_n256_k31_211", "disk_n256_k5_t28",
                                  "disk_n200_k7_wiped"])
def test_first_principles_pipeline_on_golden(name):
    g = load_golden(name)
    img = g["image"].astype(np.float64)
    n, r = img.shape[0], int(g["pixel_radius"])
    k, hb, tr = int(g["kw_medfilt_kernel"]), int(g["kw_h_bins"]), list(g["kw_threshold_range"])
    on = fp.circle_mask(n, n, int(n / 2), int(n / 2), r)
    filt = np.where(on, fp.median_zero_pad(img, k), 0.0)
    assert np.array_equal(filt, g["filt_disk"].astype(np.float64))
    cnt, edges = fp.histogram_uniform(filt[on], hb)
    st = fp.select_threshold(cnt, edges, 5)
    assert np.array_equal(st["hbase"], g["hbase"]) and np.array_equal(edges, g["hbins"])
    assert np.array_equal(np.array(st["peaks"]), g["peaks"])
    lims = fp.limits_from(st["peaks"][0], tr)
    assert np.array_equal(lims, g["limits"])
    T = len(lims)
    lev = fp.level_image(filt, on, lims)
    maps_g = np.unpackbits(g["maps"])[:T * n * n].reshape(T, n, n)
    _, maps, _, _ = fp.cleanup_levels(lev, T, r, 1e-3)
    for t in range(T):
        assert np.array_equal(maps[t], maps_g[t]), t
    pr = fp.prob_from_counts(fp.prob_counts(filt[on], st["peaks"][0], lims))
    assert np.array_equal(pr, g["probs"], equal_nan=True)
    if "stack" in g:
        K = np.full((n, n), T, dtype=np.int64)
        for t in range(T - 1, -1, -1):
            K[maps[t] == 1] = t
        stack, _ = fp.prob_stack_levels(K, pr)
        assert np.array_equal(stack, g["stack"], equal_nan=True)


@pytest.mark.parametrize("offset,scale,hb,dtype", [(0.0, 1.0, 500, np.float32), (40.0, 1.0, 5000, np.float32),
                                                   (-300.0, 0.5, 2000, np.float32), (3.0, 1e-3, 777, np.float64),
                                                   (1e6, 1.0, 200, np.float64), (0.0, 1.0, 16, np.float32)])
def test_hist_fast_path_error_bound(offset, scale, hb, dtype):
    # csrc/stats.cu hist_bin_fast / hist_eps restated in numpy float32: wherever the kernel would take
    # trunc(t) as the bin without reading an edge (t farther than eps from an integer), that bin is the
    # one whose float64 np.histogram edges bracket the value; and the actual float32 error of t stays
    # at least 8x below eps (the margin the kernel's constant claims).
    rng = np.random.default_rng(hb)
    x = np.concatenate([rng.normal(28, 3, 60000), rng.normal(130, 15, 90000), np.linspace(20, 180, 10001)])
    x = (offset + scale * x).astype(dtype)
    v = x.astype(np.float64)
    first, last = v.min(), v.max()
    _, en = np.histogram(v, bins=hb)
    true = np.clip(np.searchsorted(en, v, side="right") - 1, 0, hb - 1)
    denom = last - first
    e = 8.0 * hb * 2.0 ** -24 * (3.0 + 2.0 * max(abs(first), abs(last)) / denom)
    eps = np.float32(e) * np.float32(1.0001) if e < 0.25 else np.float32(2.0)
    t = (x.astype(np.float32) - np.float32(first)) * np.float32(hb / denom)
    fl = np.floor(t)
    fr = t - fl
    fast = (fr >= eps) & (fr <= np.float32(1.0) - eps) & (fl >= 0) & (fl < hb)
    assert np.array_equal(fl[fast].astype(np.int64), true[fast])
    err = np.max(np.abs(t.astype(np.float64) - (v - first) / denom * hb))
    assert err * 8.0 <= e
