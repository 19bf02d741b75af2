"""`Chips` — drop-in for `chips.chips.Chips` of the reference on the detection hot path.

Same constructor signature and defaults (`/root/reference/chips/chips.py:47-61`), same entry
point `run_CHIPS(wavelength, resolution, clear_prev_runs)` (`:110-150`), same stage methods
(`:173, :202, :223, :341`), same result attributes written on the caller's `disk` object
(`solar_mask`, `solar_filter`, `histogram`, `solar_ch_regions`) with the `hasattr`
memoisation and the sticky `h_thresh` of the reference.  All arithmetic runs in the sm_100a
kernels behind `libchips_b200.so`; result arrays are materialised from device memory lazily
on first attribute access.  There is no CPU fallback.

Out of scope: plotting (`plot_diagonestics`, the filament summary plot).  `run_fl_cleanup=True` uses `cleanup.CleanFl` (csrc/fl.cu).
"""
from __future__ import annotations

import os
from argparse import Namespace
from typing import List

import numpy as np
import torch

from . import engine
from .engine import Detector, Geometry

try:                                    # logging is optional
    from loguru import logger
except Exception:                       # pragma: no cover
    import logging
    logger = logging.getLogger("pychips_b200")


class LazyNamespace(Namespace):
    """`argparse.Namespace` whose listed attributes are produced on first access
    (device -> host copies happen only for what the caller actually reads)."""

    def __init__(self, _lazy=None, **kw):
        super().__init__(**kw)
        object.__setattr__(self, "_lazy", dict(_lazy or {}))

    def __getattr__(self, name):         # only called when normal lookup fails
        lazy = object.__getattribute__(self, "_lazy")
        if name in lazy:
            value = lazy.pop(name)()
            setattr(self, name, value)
            return value
        raise AttributeError(name)

    def __contains__(self, key):
        return key in self.__dict__ or key in object.__getattribute__(self, "_lazy")


class Chips(object):
    r"""CHIPS on a full-disk image, B200-native.  Parameters as in the reference."""

    def __init__(
        self,
        aia,
        base_folder: str = "tmp/chips_dataset/",
        medfilt_kernel: int = 51,
        h_bins: int = 500,
        h_thresh: float = None,
        ht_peak_ratio: int = 5,
        hist_xsplit: int = 2,
        hist_ysplit: int = 2,
        threshold_range: List[float] = [0, 20],
        porb_threshold: float = 0.8,
        area_threshold: float = 1e-3,
        run_fl_cleanup: bool = False,
        device=None,
        make_folder: bool = True,
    ) -> None:
        self.aia = aia
        self.base_folder = base_folder
        self.medfilt_kernel = medfilt_kernel
        self.h_bins = h_bins
        self.h_thresh = h_thresh
        self.ht_peak_ratio = ht_peak_ratio
        self.hist_xsplit = hist_xsplit
        self.hist_ysplit = hist_ysplit
        self.threshold_range = threshold_range
        self.porb_threshold = porb_threshold
        self.area_threshold = area_threshold
        self.run_fl_cleanup = run_fl_cleanup
        self.device = engine.require_cuda(device)
        self._det = {}                   # id(disk) -> Detector
        self._img = {}                   # id(disk) -> device image
        self._disks = {}                 # id(disk) -> disk: keeps the id alive as long as its detector
        if make_folder:
            self.get_base_folder()

    # ---- chips.py:78-88
    def get_base_folder(self) -> None:
        self.folder = self.base_folder + self.aia.date.strftime("%Y.%m.%d")
        os.makedirs(self.folder, exist_ok=True)

    # ---- chips.py:90-108
    def clear_run_results(self, disk):
        for key in ("solar_mask", "solar_filter", "histogram", "histograms", "solar_ch_regions"):
            if hasattr(disk, key):
                delattr(disk, key)
        self._det.pop(id(disk), None)
        self._img.pop(id(disk), None)
        self._disks.pop(id(disk), None)

    # ---- chips.py:110-150
    def run_CHIPS(self, wavelength: int = None, resolution: int = None,
                  clear_prev_runs: bool = False) -> None:
        if self.run_fl_cleanup:
            self.run_filament_cleanup()
        resolution = self.aia.resolution if resolution is None else resolution
        if (wavelength is not None) and (resolution is not None):
            if (wavelength in self.aia.wavelengths) and (resolution == self.aia.resolution):
                logger.info(f"Singleton: Running CHIPS for {wavelength}/{resolution}")
                disk = self.aia.datasets[wavelength][resolution]
                self.process_CHIPS(disk, clear_prev_runs)
            else:
                logger.error(f"No entry for {wavelength} and {resolution}!!!")
        else:
            for wavelength in self.aia.wavelengths:
                if (wavelength in self.aia.wavelengths) and (resolution is self.aia.resolution):
                    logger.info(f"Multiton: Running CHIPS for {wavelength}/{resolution}")
                    disk = self.aia.datasets[wavelength][resolution]
                    self.process_CHIPS(disk, clear_prev_runs)
                else:
                    logger.error(f"No entry for {wavelength} and {resolution}!!!")
        return

    # ---- chips.py:619-672.  The reference re-registers level-1 FITS files for 171/193/211 A
    # (chips/fetch.py, outside the accelerated path); here the three wavelengths must already be
    # in `self.aia.datasets`.  The summary plot (:663-670) is presentation.
    def run_filament_cleanup(self, params: dict = None, **_plot_kwargs) -> None:
        from .cleanup import DEFAULT_PARAMS, CleanFl
        logger.info("Into the filament Cleanups!")
        missing = [w for w in (171, 193, 211) if w not in self.aia.datasets]
        if missing:
            raise ValueError(
                f"run_fl_cleanup needs the 171/193/211 A disks in aia.datasets (missing {missing}); "
                "fetching / registering them is outside the accelerated path")
        self.cf = CleanFl(self.aia, resolution=self.aia.resolution,
                          params=DEFAULT_PARAMS if params is None else params, device=self.device)
        self.cf.create_coronal_hole_candidates()
        return

    # ---- chips.py:152-171 (the plotter call at :170 is presentation, not detection)
    def process_CHIPS(self, disk, clear_prev_runs) -> None:
        if clear_prev_runs:
            self.clear_run_results(disk)
        self.extract_solar_masks(disk)
        self.run_filters(disk)
        self.extract_histogram(disk)
        self.extract_CHs_CHBs(disk)
        return

    # ------------------------------------------------------------------ device plumbing
    def _geometry(self, disk) -> Geometry:
        H, W = disk.normalized.data.shape
        c = int(disk.resolution / 2)             # chips.py:188
        return Geometry(int(H), int(W), c, c, int(disk.pixel_radius))

    def _out_dtype(self, disk):
        # np.zeros_like(raw.data) * np.nan: float32 stays float32, everything else -> float64
        return np.float32 if np.asarray(disk.raw.data).dtype == np.float32 else np.float64

    def _filt_dtype(self, disk):
        # chips.py:213-216 i_mask * medfilt2d(normalized.data): the mask follows raw.data, the median
        # its input (float32 stays float32, everything else is float64)
        nd = np.asarray(disk.normalized.data).dtype
        return np.result_type(self._out_dtype(disk), np.float32 if nd == np.float32 else np.float64)

    def _detector(self, disk) -> Detector:
        det = self._det.get(id(disk))
        if det is None:
            with torch.cuda.device(self.device):
                img, elem = engine.to_device_image(np.asarray(disk.normalized.data), self.device)
                # chips.py:352: List[float], any start, unit steps.  An empty range yields no
                # regions in the reference: the stages before the thresholds still run (T = 1 plan)
                t0, t1 = self.threshold_range[0], self.threshold_range[1]
                self._no_thresholds = len(np.arange(t0, t1)) == 0
                if self._no_thresholds:
                    t0, t1 = 0, 1
                # float32 raw AND normalized arrays: h_data = filt_disk * n_mask stays float32 and numpy
                # keeps float32 bin edges (np.histogram's bin_type), bound, limit_0 and sigmoid
                f32 = (np.asarray(disk.normalized.data).dtype == np.float32
                       and np.asarray(disk.raw.data).dtype == np.float32)
                det = Detector(self._geometry(disk), self.medfilt_kernel, self.h_bins, t0, t1, elem, self.device,
                               f32_math=f32)
                det.state = set()
            self._det[id(disk)] = det
            self._img[id(disk)] = img
            self._disks[id(disk)] = disk
        return det

    def _ensure_filt(self, disk, det) -> None:
        """`disk.solar_filter` may have been memoised by an earlier `Chips` object
        (chips.py:212 `hasattr` guard): reload it instead of recomputing."""
        if "filt" not in det.state:
            host = np.ascontiguousarray(disk.solar_filter.filt_disk)
            det.filt.copy_(torch.from_numpy(host).to(self.device).to(det.dtype))
            engine._lib.check(det.lib.chips_minmax(
                engine._ptr(det.filt), det._pp, *det._g(), engine._ptr(det.ws),
                engine._stream_ptr()), "chips_minmax")
            det.state.add("filt")

    # ---- chips.py:173-200
    def extract_solar_masks(self, disk) -> None:
        logger.info(f"Create solar mask for {disk.wavelength}/{disk.resolution}")
        if not hasattr(disk, "solar_mask"):
            self.disk_area = np.pi * disk.pixel_radius ** 2
            g = self._geometry(disk)
            dt = self._out_dtype(disk)
            tdt = torch.float32 if dt == np.float32 else torch.float64

            def both():
                with torch.cuda.device(self.device):
                    n = torch.empty((g.H, g.W), dtype=tdt, device=self.device)
                    i = torch.empty((g.H, g.W), dtype=tdt, device=self.device)
                    lib = engine._lib.load()
                    engine._lib.check(lib.chips_disk_mask(
                        engine._ptr(n), engine._ptr(i), g.H, g.W, n.element_size(), g.cx, g.cy, g.r,
                        engine._stream_ptr()), "chips_disk_mask")
                    return n.cpu().numpy(), i.cpu().numpy()

            cache = {}

            def get(idx):
                if "v" not in cache:
                    cache["v"] = both()
                return cache["v"][idx]

            disk.set_value("solar_mask", LazyNamespace(
                _lazy=dict(n_mask=lambda: get(0), i_mask=lambda: get(1))))
        return

    # ---- chips.py:202-221
    def run_filters(self, disk) -> None:
        logger.info(f"Running solar filters for {disk.wavelength}/{disk.resolution}")
        if not hasattr(disk, "solar_filter"):
            det = self._detector(disk)
            img = self._img[id(disk)]
            dt = self._filt_dtype(disk)
            with torch.cuda.device(self.device):
                det.medfilt(img)
            det.state.add("filt")

            def filt_disk():
                return det.filt.cpu().numpy().astype(dt, copy=False)

            def solar_disk():
                i_mask = disk.solar_mask.i_mask
                return np.asarray(disk.normalized.data) * i_mask

            disk.set_value("solar_filter", LazyNamespace(
                _lazy=dict(solar_disk=solar_disk, filt_disk=filt_disk)))
        return

    # ---- chips.py:223-258
    def extract_histogram(self, disk) -> None:
        logger.info(f"Extract solar histogram {disk.wavelength}/{disk.resolution}")
        if not hasattr(disk, "histogram"):
            det = self._detector(disk)
            with torch.cuda.device(self.device):
                self._ensure_filt(disk, det)
                det.histogram()
                det.select_threshold(self.ht_peak_ratio, self.h_thresh)
                det.state.add("hist")
                res = det.result()
                engine._lib.check_record(res)
                if self.h_thresh is None:
                    self.h_thresh = np.float64(res.h_thresh)       # sticky, chips.py:240-241
                npk = int(res.n_peaks)
                pidx = det.peak_index[:npk].cpu().numpy()
                edt = np.float32 if det.f32_math else np.float64      # dtype of np.histogram's edges
                bound = det.bound.cpu().numpy().astype(edt, copy=False)
            peaks = [bound[p] for p in pidx]
            disk.set_value("histogram", LazyNamespace(
                _lazy=dict(hbase=lambda: det.density.cpu().numpy(),
                           hbins=lambda: det.edges.cpu().numpy().astype(edt, copy=False),
                           counts=lambda: det.counts.cpu().numpy()),
                peaks=peaks, bound=bound, h_thresh=self.h_thresh, h_bins=self.h_bins,
                peak_indexs=pidx))
        return

    # ---- chips.py:260-319.  The reference method is dead code: it is commented out of
    # process_CHIPS (:167) and raises NameError when called (`r_peaks`, `regions`, `d` are never
    # defined; its first line binds `histograms, peaks, k` instead).  This is the evident intent —
    # regions = {}, r_peaks = [], data = the region's masked pixels — with everything else as
    # written, including the column slice `y*yunit : (y+1)*xunit` and `fpeak` collecting peak
    # INDICES.  Parity unpinned (no reference output exists); pinned to oracle/chips_oracle.py.
    def extract_histograms(self, disk) -> None:
        logger.info(f"Extract solar histograms for different regions {disk.wavelength}/{disk.resolution}")
        if (not hasattr(disk, "histograms") and (self.hist_xsplit is not None)
                and (self.hist_ysplit is not None)):
            det = self._detector(disk)
            regions, r_peaks, k = {}, [], 0
            xunit, yunit = (int(disk.resolution / self.hist_xsplit), int(disk.resolution / self.hist_ysplit))
            H, W = det.geom.H, det.geom.W
            for x in range(self.hist_xsplit):
                for y in range(self.hist_ysplit):
                    y0, y1 = min(x * xunit, H), min((x + 1) * xunit, H)         # first index = rows
                    x0, x1 = min(y * yunit, W), min((y + 1) * xunit, W)
                    with torch.cuda.device(self.device):
                        self._ensure_filt(disk, det)
                        out = det.region_histogram(y0, y1, x0, max(x0, x1), self.ht_peak_ratio, self.h_thresh)
                    if out is None:        # no on-disk pixel: numpy's histogram of an empty array
                        be = np.linspace(0, 1, self.h_bins + 1)
                        h, bc, peaks, hthr = (np.full(self.h_bins, np.nan), be[:-1] + np.diff(be),
                                              np.zeros(0, dtype=np.int64), np.float64(np.nan))
                    else:
                        h, bc, peaks, hthr = out
                        bc = bc.astype(np.float32 if det.f32_math else np.float64, copy=False)
                    if self.h_thresh is None:
                        self.h_thresh = hthr
                    if len(peaks) > 1:
                        r_peaks.extend(peaks[:2])

                    def data(y0=y0, y1=y1, x0=x0, x1=x1):
                        return (disk.solar_filter.filt_disk[y0:y1, x0:x1] *
                                disk.solar_mask.n_mask[y0:y1, x0:x1]).ravel()

                    regions[k] = LazyNamespace(_lazy=dict(data=data), peaks=peaks, hbase=h, bound=bc,
                                               h_thresh=self.h_thresh, h_bins=self.h_bins)
                    k += 1
            disk.set_value("histograms", Namespace(**dict(reg=regions, fpeak=r_peaks, keys=range(k))))
        return

    # ---- chips.py:321-339: an empty stub in the reference (`pass`).  Built from its docstring
    # ("sliding a window (size determined by `xsplit`, `ysplit`) through the solar disk") and
    # docs/tutorial/workings.md:36-39: windows of the regional unit size, slid in raster order
    # with `overlap` (default one half), each processed like a region of extract_histograms
    # (chips.py:282-310).  All windows go through the device at once (chips_window_histograms:
    # one warp per window for the selection).  PARITY UNPINNED; pinned to
    # oracle/chips_oracle.py::sliding_histograms.  Sets `disk.histograms` like extract_histograms.
    def extract_sliding_histograms(self, disk, overlap: float = 0.5) -> None:
        logger.info(f"Extract solar histograms for different regions {disk.wavelength}/{disk.resolution}")
        if (not hasattr(disk, "histograms") and (self.hist_xsplit is not None)
                and (self.hist_ysplit is not None)):
            det = self._detector(disk)
            res = int(disk.resolution)
            xunit, yunit = int(res / self.hist_xsplit), int(res / self.hist_ysplit)
            sx = max(1, int(round(xunit * (1.0 - overlap))))
            sy = max(1, int(round(yunit * (1.0 - overlap))))
            windows = [(r0, r0 + xunit, c0, c0 + yunit)
                       for r0 in range(0, res - xunit + 1, sx) for c0 in range(0, res - yunit + 1, sy)]
            regions, r_peaks = {}, []
            if windows:
                with torch.cuda.device(self.device):
                    self._ensure_filt(disk, det)
                    dens, bnd, pks, hthr, _ = det.window_histograms(windows, self.ht_peak_ratio, self.h_thresh)
                    bnd = bnd.astype(np.float32 if det.f32_math else np.float64, copy=False)
                if self.h_thresh is None:
                    self.h_thresh = hthr
                for k, (r0, r1, c0, c1) in enumerate(windows):
                    if len(pks[k]) > 1:
                        r_peaks.extend(pks[k][:2])

                    def data(r0=r0, r1=r1, c0=c0, c1=c1):
                        return (disk.solar_filter.filt_disk[r0:r1, c0:c1] *
                                disk.solar_mask.n_mask[r0:r1, c0:c1]).ravel()

                    regions[k] = LazyNamespace(_lazy=dict(data=data), peaks=pks[k], hbase=dens[k], bound=bnd[k],
                                               h_thresh=self.h_thresh, h_bins=self.h_bins, window=(r0, r1, c0, c1))
            disk.set_value("histograms", Namespace(**dict(reg=regions, fpeak=r_peaks, keys=range(len(windows)))))
        return

    # ---- chips.py:341-403
    def extract_CHs_CHBs(self, disk) -> None:
        logger.info(f"Extract CHs and CHBs for different regions {disk.wavelength}/{disk.resolution}")
        if not hasattr(disk, "solar_ch_regions"):
            if len(disk.histogram.peaks) > 0:
                det = self._detector(disk)
                if self._no_thresholds:              # np.arange(...) empty: `for lim in limits` never runs
                    disk.set_value("solar_ch_regions", Namespace())
                    return
                dt = self._out_dtype(disk)
                with torch.cuda.device(self.device):
                    if "hist" not in det.state:      # results memoised on `disk` by another object
                        self._ensure_filt(disk, det)
                        det.histogram()
                        det.select_threshold(self.ht_peak_ratio, disk.histogram.h_thresh)
                        det.state.add("hist")
                    det.threshold_prob(self.porb_threshold)
                    if self.run_fl_cleanup:          # chips.py:376-377
                        engine._lib.check(det.lib.chips_apply_candidates(
                            det._pp, engine._ptr(self.cf.bits), engine._ptr(det.ws),
                            engine._stream_ptr()), "chips_apply_candidates")
                    det.cleanup(self.area_threshold)
                    det.state.add("maps")
                    res = det.result()
                engine._lib.check_record(res)
                limit_0 = (np.float32 if det.f32_math else np.float64)(res.limit_0)    # = histogram.peaks[0]
                dtmp_map = {}
                traced_all = {}
                for t in range(det.T):
                    lim = np.float64(res.limits[t])

                    def _map(t=t, which=0):
                        with torch.cuda.device(self.device):
                            return det.expand_map(t, which).cpu().numpy().astype(dt, copy=False)

                    def _ctr(which, t=t, traced=traced_all):
                        # chips.py:382-400 on demand: every border of every raw mask on the device
                        # (one enqueue for all T), then the area cut of clean_small_scale_structures
                        # (chips.py:421-429)
                        if not traced:
                            with torch.cuda.device(self.device):
                                per_t = det.contours_all()
                            for tt, (cs, hier, area2) in enumerate(per_t):
                                keep = np.flatnonzero(area2 / 2.0 / det.geom.disk_area >= self.area_threshold)
                                traced[tt] = dict(contours=[cs[i] for i in keep],
                                                  hierarchy=np.array([hier[i] for i in keep]),
                                                  contour_ids=[f"CH_{ix}" for ix in range(len(keep))])
                        return traced[t][which]

                    dtmp_map[str(lim)] = LazyNamespace(
                        _lazy=dict(map=_map, raw_map=lambda t=t: _map(t, 1),
                                   filled_map=lambda t=t: _map(t, 2),
                                   contours=lambda f=_ctr: f("contours"),
                                   hierarchy=lambda f=_ctr: f("hierarchy"),
                                   contour_ids=lambda f=_ctr: f("contour_ids")),
                        lim=lim, limit_0=limit_0, prob=np.float64(res.probs[t]),
                        n_components=int(res.n_comp[t]), n_kept=int(res.n_kept[t]), index=t)
                    logger.info(f"Estimated prob.({res.probs[t]}) at lim({lim})")
                disk.set_value("solar_ch_regions", Namespace(**dtmp_map))
        return

    # ---- plots.py:305-322 (the arithmetic only)
    def probability_map(self, disk, prob_lower_lim: float = 0.0) -> np.ndarray:
        det = self._device_results(disk)
        with torch.cuda.device(self.device):
            return det.prob_stack(prob_lower_lim).cpu().numpy()

    def labels(self, disk, lim_index: int):
        """Canonical component labels (raster order of first pixel) of the hole-filled mask
        at threshold `lim_index`, and 2*contourArea per label."""
        return self._device_results(disk).labels(lim_index)

    def _device_results(self, disk) -> Detector:
        """The detector of `disk`, with the final maps of THIS object's run resident (a disk whose
        results were memoised by another object has nothing on this object's device workspace)."""
        det = self._detector(disk)
        if "maps" not in det.state:
            raise RuntimeError("no detection of this disk has run on this Chips object: call "
                               "run_CHIPS(clear_prev_runs=True) / extract_CHs_CHBs first")
        return det

    # ---- chips.py:512-590.  The region cube is packed on the device (chips_pack_cube); the file
    # is written with netCDF4 in the reference's group layout when that package is importable,
    # otherwise as NetCDF-3 (ncio.py; the classic format has no groups: variables are named
    # "<group>_<variable>").
    def to_netcdf(self, wavelength: int, resolution: int, file_name: str = None) -> str:
        disk = self.aia.datasets[wavelength][resolution]
        file_name = (file_name if file_name
                     else f"{disk.date.strftime('%Y-%m-%d-%H-%M')}_{wavelength}A_{resolution}.nc")
        file = self.folder + f"/{file_name}"
        logger.info(f"Save to file: {file}")
        groups = {"fits": dict(fitdata=(("xarcs", "yarcs"), np.asarray(disk.normalized.data)))}
        if hasattr(disk, "solar_filter"):
            groups["filter"] = dict(
                solar_disk=(("xarcs", "yarcs"), disk.solar_filter.solar_disk),
                filt_disk=(("xarcs", "yarcs"), disk.solar_filter.filt_disk))
        L = 0
        if hasattr(disk, "solar_ch_regions"):
            cube, probs = self.region_cube(disk)
            L = len(probs)
            groups["ch_regions"] = dict(ch_regions=(("xarcs", "yarcs", "probs"), cube),
                                        probs=(("probs",), np.asarray(probs)))
        dims = dict(xarcs=resolution, yarcs=resolution, probs=L)
        try:
            from netCDF4 import Dataset
        except ImportError:
            Dataset = None
        if Dataset is not None:
            nc = Dataset(file, "w", format="NETCDF4")
            for gname, variables in groups.items():
                grp = nc.createGroup(gname)
                used = sorted({d for dd, _ in variables.values() for d in dd}, key=list(dims).index)
                for d in used:
                    grp.createDimension(d, dims[d])
                for vname, (dd, value) in variables.items():
                    grp.createVariable(vname, "f4", dd)[:] = value
            nc.close()
        else:
            from . import ncio
            flat = {f"{g}_{v}": spec for g, variables in groups.items() for v, spec in variables.items()}
            ncio.write_float32(file, {d: s for d, s in dims.items() if s}, flat)
        return file

    def region_cube(self, disk):
        """(float32 [res, res, L] cube with cube[:, :, i] = map of the i-th threshold, probs) —
        the arrays `to_netcdf` stores (chips.py:570-586)."""
        det = self._device_results(disk)
        with torch.cuda.device(self.device):
            cube = det.pack_cube(0).cpu().numpy()
        probs = [r.prob for r in disk.solar_ch_regions.__dict__.values()]
        return cube, probs

    # ---- chips.py:592-617
    def compute_similarity_measures(self, map: np.ndarray, map0: np.ndarray, measure: str = "Hu") -> dict:
        import warnings
        _, mean, used = engine.row_cosine(map, map0, self.device)
        if used == 0:
            warnings.warn("Mean of empty slice", RuntimeWarning)      # np.nanmean([]) in the reference
        return {"cosine_sim": np.float64(mean)}

    def plot_diagonestics(self, *a, **k) -> None:
        raise NotImplementedError("plotting is outside the accelerated path (SURVEY.md §2 row 6)")
