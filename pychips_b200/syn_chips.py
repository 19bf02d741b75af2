"""`SynopticChips` — drop-in for `syn_chips.SynopticChips` of the reference
(`/root/reference/chips/syn_chips.py:24-274`): same constructor defaults (`:40-52`),
`run_CHIPS()` (`:79-93`) and stage methods; no disk mask and no contour/cleanup step — the
per-threshold map is the raw threshold mask (`:237-251`).  A map with NaN/Inf inside a median
window raises `ChipsNativeError` (scipy's medfilt2d is undefined there: its selection depends on
the visiting order of the NaNs)."""
from __future__ import annotations

import os
from argparse import Namespace
from typing import List

import numpy as np
import torch

from . import engine
from .chips import LazyNamespace, logger
from .engine import Detector, Geometry


class SynopticChips(object):
    def __init__(
        self,
        synoptic_map,
        base_folder: str = "tmp/chips_dataset/",
        medfilt_kernel: int = 3,
        h_bins: int = 5000,
        h_thresh: float = None,
        ht_peak_ratio: int = 5,
        hist_xsplit: int = 2,
        hist_ysplit: int = 2,
        threshold_range: List[float] = [-5, 20],
        porb_threshold: float = 0.8,
        device=None,
        make_folder: bool = True,
    ) -> None:
        self.synoptic_map = synoptic_map
        self.base_folder = base_folder
        self.medfilt_kernel = medfilt_kernel
        self.h_bins = h_bins
        self.h_thresh = h_thresh
        self.ht_peak_ratio = ht_peak_ratio
        self.hist_xsplit = hist_xsplit
        self.hist_ysplit = hist_ysplit
        self.threshold_range = threshold_range
        self.porb_threshold = porb_threshold
        self.device = engine.require_cuda(device)
        self._det = None
        self._img = None
        if make_folder:
            self.get_base_folder()

    # ---- syn_chips.py:67-77
    def get_base_folder(self) -> None:
        self.folder = self.base_folder + self.synoptic_map.date.strftime("%Y.%m.%d")
        os.makedirs(self.folder, exist_ok=True)

    # ---- syn_chips.py:79-93 (plotter call excluded)
    def run_CHIPS(self) -> None:
        self.run_filters(self.synoptic_map)
        self.extract_histogram(self.synoptic_map)
        self.extract_CHs_CHBs(self.synoptic_map)
        return

    def _detector(self, syn) -> Detector:
        if self._det is None:
            data = np.asarray(syn.raw.data)
            with torch.cuda.device(self.device):
                img, elem = engine.to_device_image(data, self.device)
                # syn_chips.py:229: List[float], any start, unit steps; an empty range yields no regions
                t0, t1 = self.threshold_range[0], self.threshold_range[1]
                self._no_thresholds = len(np.arange(t0, t1)) == 0
                if self._no_thresholds:
                    t0, t1 = 0, 1
                H, W = data.shape
                self._det = Detector(Geometry(int(H), int(W), 0, 0, -1), self.medfilt_kernel,
                                     self.h_bins, t0, t1, elem, self.device,
                                     f32_math=data.dtype == np.float32)     # numpy's float32 histogram
                self._det.state = set()
            self._img = img
        return self._det

    def _out_dtype(self, syn):
        return np.float32 if np.asarray(syn.raw.data).dtype == np.float32 else np.float64

    # ---- syn_chips.py:95-108
    def run_filters(self, synoptic_map) -> None:
        logger.info(f"Running solar filters for {synoptic_map.wavelength}")
        if not hasattr(synoptic_map, "solar_filter"):
            det = self._detector(synoptic_map)
            with torch.cuda.device(self.device):
                det.medfilt(self._img)
                det.state.add("filt")
                filt = det.filt.cpu().numpy().astype(self._out_dtype(synoptic_map), copy=False)
            synoptic_map.set_value("solar_filter", filt)     # a bare ndarray, as in the reference
        return

    # ---- syn_chips.py:110-144
    def extract_histogram(self, synoptic_map) -> None:
        logger.info(f"Extract solar histogram {synoptic_map.wavelength}")
        if not hasattr(synoptic_map, "histogram"):
            det = self._detector(synoptic_map)
            with torch.cuda.device(self.device):
                self._histogram_on_device(synoptic_map, det, self.h_thresh)
                res = det.result()
                engine._lib.check_record(res)
                if self.h_thresh is None:
                    self.h_thresh = np.float64(res.h_thresh)
                npk = int(res.n_peaks)
                pidx = det.peak_index[:npk].cpu().numpy()
                bound = det.bound.cpu().numpy().astype(np.float32 if det.f32_math else np.float64, copy=False)
            peaks = [bound[p] for p in pidx]
            synoptic_map.set_value("histogram", LazyNamespace(
                _lazy=dict(hbase=lambda: det.density.cpu().numpy(),
                           counts=lambda: det.counts.cpu().numpy()),
                peaks=peaks, bound=bound, h_thresh=self.h_thresh, h_bins=self.h_bins,
                peak_indexs=pidx))
        return

    def _histogram_on_device(self, synoptic_map, det, h_thresh) -> None:
        """Histogram + threshold selection of `solar_filter`; a filter memoised on the map by another
        object (syn_chips.py:105 `hasattr` guard) is uploaded first."""
        if "filt" not in det.state:
            host = np.ascontiguousarray(synoptic_map.solar_filter)
            det.filt.copy_(torch.from_numpy(host).to(self.device).to(det.dtype))
            det.state.add("filt")
            det.histogram(recompute_minmax=True)
        else:
            det.histogram()
        det.select_threshold(self.ht_peak_ratio, h_thresh)
        det.state.add("hist")

    # ---- syn_chips.py:218-254
    def extract_CHs_CHBs(self, synoptic_map) -> None:
        logger.info(f"Extract CHs and CHBs for different regions {synoptic_map.wavelength}")
        if not hasattr(synoptic_map, "solar_ch_regions"):
            if len(synoptic_map.histogram.peaks) > 0:
                det = self._detector(synoptic_map)
                if self._no_thresholds:              # np.arange(...) empty: `for lim in limits` never runs
                    synoptic_map.set_value("solar_ch_regions", Namespace())
                    return
                dt = self._out_dtype(synoptic_map)
                with torch.cuda.device(self.device):
                    if "hist" not in det.state:      # results memoised on the map by another object
                        self._histogram_on_device(synoptic_map, det, synoptic_map.histogram.h_thresh)
                    det.threshold_prob(self.porb_threshold)
                    det.cleanup(0.0)
                    det.state.add("maps")
                    res = det.result()
                engine._lib.check_record(res)
                limit_0 = (np.float32 if det.f32_math else np.float64)(res.limit_0)
                dtmp_map = {}
                for t in range(det.T):
                    lim = np.float64(res.limits[t])

                    def _map(t=t):
                        with torch.cuda.device(self.device):
                            return det.expand_map(t, 1).cpu().numpy().astype(dt, copy=False)

                    dtmp_map[str(lim)] = LazyNamespace(
                        _lazy=dict(map=_map), lim=lim, limit_0=limit_0,
                        prob=np.float64(res.probs[t]), index=t)
                    logger.info(f"Estimated prob.({res.probs[t]}) at lim({lim})")
                synoptic_map.set_value("solar_ch_regions", Namespace(**dtmp_map))
        return

    # ---- BASELINE config 4 "wrap-around CCL" (no counterpart in syn_chips.py: the reference's
    # synoptic map is the raw threshold mask, :237-251).  Extension, parity unpinned.
    def labels(self, lim_index: int, periodic: bool = True):
        """Connected regions (8-connectivity, longitude periodic) of the map of threshold
        `lim_index`: (labels int32 [H, W], 0 = background, numbered in raster order; pixels per label)."""
        if self._det is None or "maps" not in self._det.state:
            raise RuntimeError("run_CHIPS() / extract_CHs_CHBs() first")
        with torch.cuda.device(self.device):
            return self._det.labels_periodic(lim_index, periodic)

    # ---- plots.py:624-641 (the arithmetic only)
    def probability_map(self, prob_lower_lim: float = 0.0) -> np.ndarray:
        with torch.cuda.device(self.device):
            return self._det.prob_stack(prob_lower_lim).cpu().numpy()

    def plot_diagonestics(self, *a, **k) -> None:
        raise NotImplementedError("plotting is outside the accelerated path (SURVEY.md §2 row 6)")
