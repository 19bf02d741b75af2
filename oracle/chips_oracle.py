"""TEST INFRASTRUCTURE — CPU oracle for the CHIPS detection hot path.  NOT product code.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; nothing under `pychips_b200/` does.

What this is
------------
A restatement, stage by stage, of the reference's hot path
(`/root/reference/chips/chips.py:173-450`, `/root/reference/chips/syn_chips.py:95-274`,
`/root/reference/chips/plots.py:305-322`).  The reference's arithmetic lives in third-party
libraries that are NOT vendored under `/root/reference` — `scipy.signal.medfilt2d`,
`scipy.signal.find_peaks`, `numpy.histogram/round/exp/linspace`, `cv2.circle`,
`cv2.findContours/contourArea/drawContours` (pinned by the reference's conda env at
numpy 1.19.2 / scipy 1.6.0 / opencv-python 4.5.1.48, `chips.yml:90,119,166`; this image has
numpy 2.3 / scipy 1.18 / OpenCV 4.13).  The oracle therefore calls the same library entry
points at the same call sites, in the same order, with the same dtypes — it is "reference
control flow + container library versions", which is exactly what parity is measured
against.

Pinning
-------
The reference ships no golden vectors and no numeric assertions.  This
oracle is pinned instead against outputs of the UNMODIFIED reference classes run in the
build container (`oracle/ref_harness.py`): `tests/golden/make_golden.py` runs both on the
same seeded inputs, asserts equality of every stage output, and commits digests + small
arrays under `tests/golden/`; `tests/test_oracle_golden.py` re-checks the oracle against
those fixtures wherever the tests run (the reference tree does not travel to the GPU box).
"""
from __future__ import annotations

from argparse import Namespace

import cv2
import numpy as np
from scipy import signal


# --------------------------------------------------------------------------- row C
def solar_masks(raw_like: np.ndarray, resolution: int, pixel_radius: int):
    """chips.py:173-200.  NaN canvas, filled cv2.circle(255), >1 -> 1; i_mask = NaN -> 0."""
    n_mask = np.zeros_like(raw_like) * np.nan
    cv2.circle(n_mask, (int(resolution / 2), int(resolution / 2)), pixel_radius, 255, -1)
    n_mask[n_mask > 1] = 1.0
    i_mask = np.copy(n_mask)
    i_mask[np.isnan(i_mask)] = 0.0
    i_mask[i_mask > 1] = 1.0
    return n_mask, i_mask


# --------------------------------------------------------------------------- row D
def run_filters(normalized: np.ndarray, i_mask: np.ndarray, medfilt_kernel: int):
    """chips.py:202-221: median of the UNMASKED image, mask applied afterwards."""
    solar_disk = normalized * i_mask
    filt_disk = i_mask * signal.medfilt2d(normalized, medfilt_kernel)
    return solar_disk, filt_disk


def syn_run_filters(raw: np.ndarray, medfilt_kernel: int):
    """syn_chips.py:95-108."""
    return signal.medfilt2d(raw, medfilt_kernel)


# --------------------------------------------------------------------------- row E
def extract_histogram(h_data: np.ndarray, h_bins: int, h_thresh, ht_peak_ratio):
    """chips.py:223-258 / syn_chips.py:110-144.  `h_data` already carries NaN off-disk.
    Returns the Namespace fields plus the (possibly newly set) sticky h_thresh."""
    flat = h_data.ravel()
    h, be = np.histogram(flat[~np.isnan(flat)], bins=h_bins, density=True)
    if h_thresh is None:
        h_thresh = np.max(h) / ht_peak_ratio
    peak_indexs, _ = signal.find_peaks(h, height=h_thresh)
    bc = be[:-1] + np.diff(be)
    peaks = [bc[p] for p in peak_indexs]
    return dict(peaks=peaks, hbase=h, hbins=be, bound=bc, h_thresh=h_thresh, h_bins=h_bins,
                peak_indexs=np.asarray(peak_indexs))


# --------------------------------------------------------------------------- row G
def calculate_prob(data: np.ndarray, thresholds, porb_threshold: float = 0.8):
    """chips.py:432-450 / syn_chips.py:256-274."""
    data = data[~np.isnan(data)]
    data = data[data <= thresholds[0]]
    ps = 1.0 / (1.0 + np.exp(data - thresholds[1]).round(4))
    h, e = np.histogram(ps, bins=np.linspace(0, 1, 21))
    b = (e + np.diff(e)[0] / 2)[:-1]
    with np.errstate(invalid="ignore", divide="ignore"):
        p = np.sum(b[b > porb_threshold] * h[b > porb_threshold]) / np.sum(b * h)
    return np.round(p, 4)


# --------------------------------------------------------------------------- row H
def clean_small_scale_structures(contours, hierarchy, data, disk_area, area_threshold):
    """chips.py:405-430."""
    dxmap = np.copy(data)
    new_contours, new_hierarchy = list(), list()
    for i, cnt in enumerate(contours):
        area = cv2.contourArea(cnt)
        if area / disk_area >= area_threshold:
            new_contours.append(cnt)
            if hierarchy[0][i][3] < 0:
                cv2.drawContours(dxmap, [cnt], -1, (255, 255, 255), -1)
            new_hierarchy.append(hierarchy[0][i])
    return new_contours, np.array(new_hierarchy), dxmap / 255


# --------------------------------------------------------------------------- rows F, I
def extract_CHs_CHBs(filt_disk, n_mask, resolution, pixel_radius, peaks, threshold_range,
                     porb_threshold=0.8, area_threshold=1e-3, keep_contours=True,
                     candidate_bitmap=None):
    """chips.py:341-403.  Returns an ordered dict str(lim) -> Namespace, or None when the
    histogram had no peak (chips.py:355: `solar_ch_regions` is then never set)."""
    tr = np.arange(threshold_range[0], threshold_range[1])
    if len(peaks) == 0:
        return None
    disk_area = np.pi * pixel_radius ** 2
    limit_0 = peaks[0]
    limits = np.round(limit_0 + tr, 4)
    data = filt_disk * n_mask
    out = {}
    for lim in limits:
        tmp_data = np.copy(data)
        nd_mask = np.zeros_like(data) * np.nan
        cv2.circle(nd_mask, (int(resolution / 2), int(resolution / 2)), pixel_radius, 255, -1)
        nd_mask[nd_mask > 0] = 1
        tmp_data = nd_mask * tmp_data
        with np.errstate(invalid="ignore"):
            tmp_data[tmp_data <= lim] = -1
            tmp_data[tmp_data > lim] = 0
            tmp_data[tmp_data == -1] = 1
        tmp_data[np.isnan(tmp_data)] = 0
        if candidate_bitmap is not None:                       # chips.py:376-377 (run_fl_cleanup)
            tmp_data = tmp_data * candidate_bitmap
        p = calculate_prob(np.copy(data).ravel(), [lim, limit_0], porb_threshold)
        contours, hierarchy = cv2.findContours(tmp_data.astype(np.uint8), cv2.RETR_TREE,
                                               cv2.CHAIN_APPROX_SIMPLE)
        contours, hierarchy, dxmap = clean_small_scale_structures(
            contours, hierarchy, np.zeros_like(tmp_data), disk_area, area_threshold)
        out[str(lim)] = Namespace(lim=lim, limit_0=limit_0, prob=p, map=dxmap,
                                  raw_map=tmp_data,
                                  contours=contours if keep_contours else None,
                                  hierarchy=hierarchy if keep_contours else None,
                                  contour_ids=[f"CH_{ix}" for ix in range(len(contours))])
    return out


def syn_extract_CHs_CHBs(solar_filter, peaks, threshold_range, porb_threshold=0.8):
    """syn_chips.py:218-254 (no mask, no contour step)."""
    tr = np.arange(threshold_range[0], threshold_range[1])
    if len(peaks) == 0:
        return None
    limit_0 = peaks[0]
    limits = np.round(limit_0 + tr, 4)
    data = solar_filter
    out = {}
    for lim in limits:
        tmp_data = np.copy(data)
        with np.errstate(invalid="ignore"):
            tmp_data[tmp_data <= lim] = -1
            tmp_data[tmp_data > lim] = 0
            tmp_data[tmp_data == -1] = 1
        tmp_data[np.isnan(tmp_data)] = 0
        p = calculate_prob(np.copy(data).ravel(), [lim, limit_0], porb_threshold)
        out[str(lim)] = Namespace(lim=lim, limit_0=limit_0, prob=p, map=np.copy(tmp_data))
    return out


# --------------------------------------------------------------------------- row J
def prob_stack(maps, probs, prob_lower_lim: float = 0.0):
    """plots.py:305-322 / :624-641: max_i(map_i * n_prob_i), 0 -> NaN."""
    probs = np.array(probs)
    with np.errstate(invalid="ignore", divide="ignore"):
        n_probs = (probs - probs.min()) / (probs.max() - probs.min())
    stacked = np.max([m * p for m, p in zip(maps, n_probs) if p >= prob_lower_lim], axis=0)
    stacked[stacked == 0.0] = np.nan
    return stacked, n_probs


# --------------------------------------------------------------------------- rows A, B
class OracleChips:
    """Stage order of `Chips.process_CHIPS` (chips.py:152-171) without the plotter, with the
    reference's sticky `h_thresh` (chips.py:240-241) and `hasattr` memoisation."""

    def __init__(self, medfilt_kernel=51, h_bins=500, h_thresh=None, ht_peak_ratio=5,
                 threshold_range=(0, 20), porb_threshold=0.8, area_threshold=1e-3,
                 candidate_bitmap=None):
        self.candidate_bitmap = candidate_bitmap     # CleanFl output when run_fl_cleanup=True
        self.medfilt_kernel = medfilt_kernel
        self.h_bins = h_bins
        self.h_thresh = h_thresh
        self.ht_peak_ratio = ht_peak_ratio
        self.threshold_range = list(threshold_range)
        self.porb_threshold = porb_threshold
        self.area_threshold = area_threshold

    def run(self, disk, keep_contours=True):
        n_mask, i_mask = solar_masks(disk.raw.data, disk.resolution, disk.pixel_radius)
        disk.set_value("solar_mask", Namespace(n_mask=n_mask, i_mask=i_mask))
        solar_disk, filt_disk = run_filters(disk.normalized.data, i_mask, self.medfilt_kernel)
        disk.set_value("solar_filter", Namespace(solar_disk=solar_disk, filt_disk=filt_disk))
        hd = extract_histogram(filt_disk * n_mask, self.h_bins, self.h_thresh,
                               self.ht_peak_ratio)
        self.h_thresh = hd["h_thresh"]
        disk.set_value("histogram", Namespace(**hd))
        regs = extract_CHs_CHBs(filt_disk, n_mask, disk.resolution, disk.pixel_radius,
                                hd["peaks"], self.threshold_range, self.porb_threshold,
                                self.area_threshold, keep_contours, self.candidate_bitmap)
        if regs is not None:
            disk.set_value("solar_ch_regions", Namespace(**regs))
        return disk


class OracleSynopticChips:
    """Stage order of `SynopticChips.run_CHIPS` (syn_chips.py:79-93) without the plotter."""

    def __init__(self, medfilt_kernel=3, h_bins=5000, h_thresh=None, ht_peak_ratio=5,
                 threshold_range=(-5, 20), porb_threshold=0.8):
        self.medfilt_kernel = medfilt_kernel
        self.h_bins = h_bins
        self.h_thresh = h_thresh
        self.ht_peak_ratio = ht_peak_ratio
        self.threshold_range = list(threshold_range)
        self.porb_threshold = porb_threshold

    def run(self, syn):
        filt = syn_run_filters(syn.raw.data, self.medfilt_kernel)
        syn.set_value("solar_filter", filt)
        hd = extract_histogram(filt, self.h_bins, self.h_thresh, self.ht_peak_ratio)
        self.h_thresh = hd["h_thresh"]
        hd_ns = dict(hd)
        hd_ns.pop("hbins")               # syn_chips.py:131-141 stores no `hbins`
        syn.set_value("histogram", Namespace(**hd_ns))
        regs = syn_extract_CHs_CHBs(filt, hd["peaks"], self.threshold_range,
                                    self.porb_threshold)
        if regs is not None:
            syn.set_value("solar_ch_regions", Namespace(**regs))
        return syn


# --------------------------------------------------------------------------- SURVEY §8(f) rows 2, 4
def netcdf_arrays(disk):
    """The arrays `Chips.to_netcdf` stores in the `ch_regions` group (chips.py:570-586):
    `data[:, :, i] = region.map` as float32 ("f4" variable) and the list of probabilities."""
    keys = list(disk.solar_ch_regions.__dict__.keys())
    res = disk.resolution
    data, probs = np.zeros((res, res, len(keys))), []
    for i, k in enumerate(keys):
        region = getattr(disk.solar_ch_regions, str(k))
        data[:, :, i] = region.map
        probs.append(region.prob)
    return data.astype(np.float32), probs


def compute_similarity_measures(map, map0):
    """chips.py:592-617 (row-wise cosine similarity through scikit-learn, then np.nanmean)."""
    from sklearn.metrics.pairwise import cosine_similarity
    measures = {"cosine_sim": np.nan}
    cosine_sim = []
    for i in range(map.shape[0]):
        if np.sum(map[i, :]) > 0:
            a, b = (np.array([list(map[i, :])]), np.array([list(map0[i, :])]))
            cosine_sim.append(cosine_similarity(a, b)[0, 0])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        measures["cosine_sim"] = np.nanmean(cosine_sim)
    return measures


# --------------------------------------------------------------------------- SURVEY §8(f) row 3
def extract_histograms(filt_disk, n_mask, resolution, hist_xsplit, hist_ysplit, h_bins, h_thresh,
                       ht_peak_ratio):
    """Regional histograms, chips.py:260-319.  PARITY UNPINNED: the reference method is dead code
    (commented out at chips.py:167) and raises NameError when called — `r_peaks`, `regions` and
    `d` are never defined (its first line binds `histograms, peaks, k`).  This restatement makes
    the three evident repairs (regions = {}, r_peaks = [], data = the region's masked pixels) and
    keeps everything else as written, including the column slice `y*yunit : (y+1)*xunit` and
    `fpeak` collecting peak indices.  Returns (regions dict, fpeak list, sticky h_thresh)."""
    regions, r_peaks, k = {}, [], 0
    xunit, yunit = int(resolution / hist_xsplit), int(resolution / hist_ysplit)
    for x in range(hist_xsplit):
        for y in range(hist_ysplit):
            r_disk_data = (filt_disk[x * xunit:(x + 1) * xunit, y * yunit:(y + 1) * xunit] *
                           n_mask[x * xunit:(x + 1) * xunit, y * yunit:(y + 1) * xunit]).ravel()
            with np.errstate(invalid="ignore", divide="ignore"):
                h, be = np.histogram(r_disk_data[~np.isnan(r_disk_data.ravel())], bins=h_bins, density=True)
            if h_thresh is None:
                h_thresh = np.max(h) / ht_peak_ratio
            peaks, _ = signal.find_peaks(h, height=h_thresh)
            bc = be[:-1] + np.diff(be)
            if len(peaks) > 1:
                r_peaks.extend(peaks[:2])
            regions[k] = Namespace(peaks=peaks, hbase=h, bound=bc, data=r_disk_data, h_thresh=h_thresh,
                                   h_bins=h_bins)
            k += 1
    return regions, r_peaks, h_thresh


def sliding_windows(resolution, hist_xsplit, hist_ysplit, overlap=0.5):
    """Window rectangles (row0, row1, col0, col1) of `extract_sliding_histograms`: windows of
    (resolution/hist_xsplit) rows x (resolution/hist_ysplit) columns - the units of
    chips.py:276-279, first index = rows as in :282-289 - slid in raster order with a step of
    (1 - overlap) of the window (at least one pixel), every window fully inside the image."""
    xunit, yunit = int(resolution / hist_xsplit), int(resolution / hist_ysplit)
    sx, sy = max(1, int(round(xunit * (1.0 - overlap)))), max(1, int(round(yunit * (1.0 - overlap))))
    out = []
    for r0 in range(0, resolution - xunit + 1, sx):
        for c0 in range(0, resolution - yunit + 1, sy):
            out.append((r0, r0 + xunit, c0, c0 + yunit))
    return out


def sliding_histograms(filt_disk, n_mask, resolution, hist_xsplit, hist_ysplit, h_bins, h_thresh,
                       ht_peak_ratio, overlap=0.5):
    """`Chips.extract_sliding_histograms`, chips.py:321-339.  PARITY UNPINNED: the reference method
    is an empty stub (`pass`); its docstring ("sliding a window (size determined by xsplit, ysplit)
    through the solar disk") and docs/tutorial/workings.md:36-39 give the intent.  This restatement
    applies the loop body of `extract_histograms` (chips.py:282-310, with the repairs described
    there) to every window of `sliding_windows`.  Returns (regions dict, fpeak list, sticky h_thresh)."""
    regions, r_peaks, k = {}, [], 0
    for (r0, r1, c0, c1) in sliding_windows(resolution, hist_xsplit, hist_ysplit, overlap):
        r_disk_data = (filt_disk[r0:r1, c0:c1] * n_mask[r0:r1, c0:c1]).ravel()
        with np.errstate(invalid="ignore", divide="ignore"):
            h, be = np.histogram(r_disk_data[~np.isnan(r_disk_data.ravel())], bins=h_bins, density=True)
        if h_thresh is None:
            h_thresh = np.max(h) / ht_peak_ratio
        peaks, _ = signal.find_peaks(h, height=h_thresh)
        bc = be[:-1] + np.diff(be)
        if len(peaks) > 1:
            r_peaks.extend(peaks[:2])
        regions[k] = Namespace(peaks=peaks, hbase=h, bound=bc, data=r_disk_data, h_thresh=h_thresh,
                               h_bins=h_bins, window=(r0, r1, c0, c1))
        k += 1
    return regions, r_peaks, h_thresh
