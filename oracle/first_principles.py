"""TEST INFRASTRUCTURE — first-principles CPU restatements.  NOT product code.

`oracle/chips_oracle.py` restates the reference by calling the same third-party entry
points it calls.  This module restates what those third-party calls *compute*, in plain
numpy, in exactly the formulation the CUDA kernels use — it is the executable
specification of each kernel and is itself checked against scipy / numpy / OpenCV in
`tests/test_first_principles.py`:

* `circle_mask`        == filled `cv2.circle` (chips.py:186-192)         integer test
* `median_zero_pad`    == `scipy.signal.medfilt2d` (chips.py:214-216)    order statistic
* `histogram_uniform`  == `np.histogram(bins=int)` (chips.py:235-239)    fp64 index + +-1 fix
* `find_peaks_height`  == `scipy.signal.find_peaks(height=)` (chips.py:242)
* `prob_bin_cuts` / `prob_from_counts` == `calculate_prob` (chips.py:432-450)
* `level_image`, `fill_levels`, `cleanup_levels` == thresholding + findContours /
  contourArea / drawContours (chips.py:360-389, 405-430) as: per-pixel first-threshold
  level, greyscale hole filling (max-min path to the border, 4-connected), 8-connected
  components of each level set, bit-quad area 2*Q4+Q3, area filter.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage


# ----------------------------------------------------------------------------- circle
def circle_mask(h: int, w: int, cx: int, cy: int, r: int) -> np.ndarray:
    """bool [h,w]: (x-cx)^2 + (y-cy)^2 <= r^2 in integers."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    return (xx - cx) ** 2 + (yy - cy) ** 2 <= int(r) * int(r)


# ----------------------------------------------------------------------------- median
def median_zero_pad(img: np.ndarray, k: int) -> np.ndarray:
    """Element of rank (k*k-1)/2 of each k x k window, zero padding.  Small images only."""
    h, w = img.shape
    p = k // 2
    pad = np.zeros((h + 2 * p, w + 2 * p), dtype=img.dtype)
    pad[p:p + h, p:p + w] = img
    win = np.lib.stride_tricks.sliding_window_view(pad, (k, k)).reshape(h, w, k * k)
    return np.partition(win, (k * k - 1) // 2, axis=-1)[..., (k * k - 1) // 2].copy()


# ----------------------------------------------------------------------------- histogram
def linspace_edges(first: float, last: float, nbins: int) -> np.ndarray:
    """np.linspace(first,last,nbins+1) in fp64: i*step + first (mul THEN add, no fma),
    last element forced to `last`."""
    first = np.float64(first)
    last = np.float64(last)
    step = (last - first) / np.float64(nbins)
    y = np.arange(0, nbins + 1, dtype=np.float64)
    if step == 0:
        y = y / np.float64(nbins)
        y = y * (last - first)
    else:
        y = y * step
    y = y + first
    y[-1] = last
    return y


def histogram_uniform(x: np.ndarray, nbins: int):
    """Counts (int64) and edges (fp64) as np.histogram(x, bins=nbins) computes them."""
    x = np.asarray(x, dtype=np.float64)
    if x.size == 0:
        first, last = np.float64(0), np.float64(1)
    else:
        first, last = x.min(), x.max()
    if first == last:
        first, last = first - 0.5, last + 0.5
    edges = linspace_edges(first, last, nbins)
    denom = last - first
    f = ((x - first) / denom) * nbins
    idx = f.astype(np.int64)
    idx[idx == nbins] -= 1
    idx[x < edges[idx]] -= 1
    inc = (x >= edges[idx + 1]) & (idx != nbins - 1)
    idx[inc] += 1
    return np.bincount(idx, minlength=nbins).astype(np.int64), edges


def density_from_counts(counts: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """`n / db / n.sum()` (numpy `_histograms_impl.histogram`, density=True)."""
    db = np.diff(edges)
    return counts / db / counts.sum()


def find_peaks_height(x: np.ndarray, height: float) -> np.ndarray:
    """Strict local maxima; a plateau yields its (left+right)//2 sample; first and last
    sample are never peaks; keep x[p] >= height."""
    n = len(x)
    peaks = []
    i = 1
    imax = n - 1
    while i < imax:
        if x[i - 1] < x[i]:
            ia = i + 1
            while ia < imax and x[ia] == x[i]:
                ia += 1
            if x[ia] < x[i]:
                peaks.append((i + ia - 1) // 2)
                i = ia
        i += 1
    peaks = np.array(peaks, dtype=np.int64)
    return peaks[x[peaks] >= height] if len(peaks) else peaks


def select_threshold(counts, edges, ht_peak_ratio, h_thresh=None):
    """chips.py:235-244 after the counts: density, sticky h_thresh, peaks, bound."""
    h = density_from_counts(counts, edges)
    if h_thresh is None:
        h_thresh = np.max(h) / ht_peak_ratio
    pk = find_peaks_height(h, h_thresh)
    bc = edges[:-1] + np.diff(edges)
    return dict(hbase=h, h_thresh=h_thresh, peak_indexs=pk, bound=bc,
                peaks=[bc[p] for p in pk])


def limits_from(limit_0, threshold_range):
    """`np.round(limit_0 + arange(t0,t1), 4)` == rint(x*1e4)/1e4 (chips.py:357)."""
    x = np.float64(limit_0) + np.arange(threshold_range[0], threshold_range[1])
    return np.rint(x * 1e4) / 1e4


# ----------------------------------------------------------------------------- probability
PROB_EDGES = np.linspace(0, 1, 21)
PROB_CENTRES = (PROB_EDGES + np.diff(PROB_EDGES)[0] / 2)[:-1]


def prob_bin_of_k(kint: np.ndarray) -> np.ndarray:
    """20-bin index of ps = 1/(1 + k/1e4) for integer-valued k = rint(exp(d)*1e4)
    (np.histogram with explicit edges: right-open bins, last bin closed)."""
    ps = 1.0 / (1.0 + np.asarray(kint, dtype=np.float64) / 1e4)
    b = np.searchsorted(PROB_EDGES, ps, side="right") - 1
    return np.clip(b, 0, 19)


def prob_k_cuts() -> np.ndarray:
    """cut[j] (j=1..19) = largest integer k whose ps still falls in bin >= j."""
    cuts = np.zeros(20, dtype=np.int64)
    for j in range(1, 20):
        lo, hi = 0, 10 ** 7           # bin(lo) >= j (k=0 -> ps=1 -> bin 19); bin(hi) = 0
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if prob_bin_of_k(np.array([mid]))[0] >= j:
                lo = mid
            else:
                hi = mid
        cuts[j] = lo
    return cuts


def prob_d_cuts() -> np.ndarray:
    """cutd[j] (j=1..19) = largest float64 d with rint(np.exp(d)*1e4) <= k_cut[j]; so
    bin(d) = #{j in 1..19 : d <= cutd[j]}.  Found by bisection on the float64 lattice with
    numpy's own exp, i.e. the constants embed the oracle's exp near the 19 cut points."""
    kc = prob_k_cuts()
    out = np.zeros(20, dtype=np.float64)
    for j in range(1, 20):
        lo, hi = np.float64(-50.0), np.float64(50.0)    # k(lo)=0 <= cut ; k(hi) huge
        lo_i, hi_i = lo.view(np.int64), hi.view(np.int64)
        # order-preserving integer view of doubles
        def key(v):
            i = np.float64(v).view(np.int64)
            return int(i) if i >= 0 else int(-(i & 0x7FFFFFFFFFFFFFFF))
        def unkey(i):
            if i >= 0:
                return np.int64(i).view(np.float64)
            return np.int64(-i | -0x8000000000000000).view(np.float64)
        a, b = key(lo), key(hi)
        while b - a > 1:
            m = (a + b) // 2
            d = unkey(m)
            if np.rint(np.exp(d) * 1e4) <= kc[j]:
                a = m
            else:
                b = m
        out[j] = unkey(a)
    return out


def prob_counts(values: np.ndarray, limit_0: float, limits: np.ndarray) -> np.ndarray:
    """[T+1, 20] table: row l = pixels whose first satisfied threshold is l (values above
    every limit are not counted); bin from d = v - limit_0."""
    v = np.asarray(values, dtype=np.float64)
    lev = np.searchsorted(np.asarray(limits, dtype=np.float64), v, side="left")
    # level = #{t: lim_t < v}
    k = np.rint(np.exp(v - np.float64(limit_0)) * 1e4)
    b = prob_bin_of_k(k)
    tab = np.zeros((len(limits) + 1, 20), dtype=np.int64)
    np.add.at(tab, (lev, b), 1)
    tab[len(limits)] = 0
    return tab


def _pairwise_sum(a):
    """numpy's pairwise add.reduce for n < 128 float64 (8 accumulators)."""
    n = len(a)
    if n < 8:
        res = np.float64(0.0)
        for v in a:
            res = res + v
        return res
    r = [np.float64(a[j]) for j in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] = r[j] + a[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res = res + a[i]
        i += 1
    return res


def prob_from_counts(tab: np.ndarray, porb_threshold: float = 0.8) -> np.ndarray:
    """T probabilities from the [T+1,20] table (prefix over levels), chips.py:445-450."""
    T = tab.shape[0] - 1
    cum = np.cumsum(tab[:T], axis=0)
    out = np.zeros(T)
    b = PROB_CENTRES
    sel = b > porb_threshold
    for t in range(T):
        h = cum[t]
        with np.errstate(invalid="ignore", divide="ignore"):
            p = _pairwise_sum(b[sel] * h[sel]) / _pairwise_sum(b * h)
        out[t] = np.round(p, 4)
    return out


# ----------------------------------------------------------------------------- levels / CCL
def level_image(filt: np.ndarray, on_disk: np.ndarray | None, limits: np.ndarray) -> np.ndarray:
    """uint8 [H,W]: first threshold index t with v <= lim_t and lim_t >= -1 (the reference
    wipes a mask whose lim < -1, chips.py:372-374); T where none (or off-disk)."""
    lim = np.asarray(limits, dtype=np.float64)
    T = len(lim)
    lev = np.searchsorted(lim, filt.astype(np.float64), side="left")
    t_inv = int(np.sum(lim < -1.0))            # limits ascend: invalid ones are a prefix
    lev = np.maximum(lev, t_inv)
    lev = np.minimum(lev, T)
    if on_disk is not None:
        lev = np.where(on_disk, lev, T)
    return lev.astype(np.uint8)


def fill_levels(lev: np.ndarray, T: int) -> np.ndarray:
    """Greyscale hole filling: W(p) = max over 4-connected paths p -> outside the image of the
    min level along the path (outside = T).  {W <= t} is the complement of the outer
    background of mask_t = {lev <= t}, i.e. mask_t with every hole (and anything nested in
    it) filled."""
    h, w = lev.shape
    W = np.zeros((h + 2, w + 2), dtype=np.int32)
    L = np.full((h + 2, w + 2), T, dtype=np.int32)
    L[1:-1, 1:-1] = lev
    W[0, :] = W[-1, :] = W[:, 0] = W[:, -1] = T
    four = ndimage.generate_binary_structure(2, 1)
    # per level, descending: outer background of {lev > t} gets W >= t+1
    for t in range(T - 1, -1, -1):
        bg = L > t
        lab, _ = ndimage.label(bg, structure=four)
        outer = lab == lab[0, 0]
        W[outer & (W < t + 1)] = t + 1
    return W[1:-1, 1:-1].astype(np.uint8)


def twice_area_quads(filled: np.ndarray, labels: np.ndarray, nlab: int) -> np.ndarray:
    """2*area of the outer contour of each 8-connected component of a hole-free set:
    2*Q4 + Q3 over all 2x2 pixel quads of the zero-padded image (polygon through pixel
    centres).  A quad with >= 3 set pixels carries a single label."""
    f = np.pad(filled.astype(np.int32), 1)
    l = np.pad(labels.astype(np.int64), 1)
    s = f[:-1, :-1] + f[:-1, 1:] + f[1:, :-1] + f[1:, 1:]
    lq = np.maximum(np.maximum(l[:-1, :-1], l[:-1, 1:]), np.maximum(l[1:, :-1], l[1:, 1:]))
    a2 = np.zeros(nlab + 1, dtype=np.int64)
    np.add.at(a2, lq[s == 4], 2)
    np.add.at(a2, lq[s == 3], 1)
    return a2


def cleanup_levels(lev: np.ndarray, T: int, pixel_radius: int, area_threshold: float):
    """Final maps for all T thresholds from the level image, plus per-threshold label image
    (raster order of first pixel) and 2*area table."""
    W = fill_levels(lev, T)
    eight = np.ones((3, 3), dtype=bool)
    disk_area = np.pi * pixel_radius ** 2
    maps, labels, areas = [], [], []
    for t in range(T):
        filled = W <= t
        lab, n = ndimage.label(filled, structure=eight)
        a2 = twice_area_quads(filled, lab, n)
        keep = (a2 * 0.5) / disk_area >= area_threshold
        keep[0] = False
        maps.append(keep[lab].astype(np.uint8))
        labels.append(lab)
        areas.append(a2)
    return W, maps, labels, areas


# ----------------------------------------------------------------------------- prob stack
def prob_stack_levels(K: np.ndarray, probs: np.ndarray, lower_lim: float = 0.0):
    """plots.py:305-322 from the kept-level image K (first t at which the pixel is in the
    final map; T if never): max over t >= K(p) of n_prob_t (with n_prob_t >= lower_lim)."""
    probs = np.asarray(probs, dtype=np.float64)
    T = len(probs)
    with np.errstate(invalid="ignore", divide="ignore"):
        n_probs = (probs - probs.min()) / (probs.max() - probs.min())
    if np.isnan(n_probs).any():
        # a NaN anywhere poisons np.max over the stack for every pixel
        return np.full(K.shape, np.nan), n_probs
    use = np.where(n_probs >= lower_lim, n_probs, 0.0)
    suffix = np.maximum.accumulate(use[::-1])[::-1]          # max_{t' >= t}
    suffix = np.append(suffix, 0.0)
    out = suffix[np.minimum(K, T)]
    out[out == 0.0] = np.nan
    return out, n_probs
