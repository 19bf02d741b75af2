"""Device-side driver of one CHIPS detection: plan + workspace + C-ABI calls.

PyTorch is used only as plumbing here (device memory, streams, host<->device copies); every
compute step is a hand-written sm_100a kernel reached through `libchips_b200.so`.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import Plan, Result

_NP2T = {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64}


def _stream_ptr(stream=None) -> C.c_void_p:
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.ChipsNativeError(
            "pychips_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


@dataclass
class Geometry:
    H: int
    W: int
    cx: int
    cy: int
    r: int            # < 0: unmasked (synoptic)

    @property
    def disk_area(self) -> float:
        return float(np.pi * self.r ** 2)      # chips.py:185


class Detector:
    """Plan + workspace for one (geometry, k, h_bins, T, elem) and the stage calls.

    Stage methods mirror the C ABI one to one; `detect` is the fused sequence
    (`chips_detect`).  All calls only enqueue on the current torch stream.
    """

    def __init__(self, geom: Geometry, k: int, h_bins: int, t0, t1, elem: int = 4,
                 device=None, ws: torch.Tensor | None = None, f32_math: bool = False):
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.geom = geom
        self.k, self.h_bins, self.t0, self.t1, self.elem = int(k), int(h_bins), t0, t1, elem
        # chips.py:352 np.arange(threshold_range[0], threshold_range[1]): List[float] in the
        # reference, any start, unit steps; the offsets go to the device as they are
        self.offsets = np.ascontiguousarray(np.arange(t0, t1), dtype=np.float64)
        self._offp = self.offsets.ctypes.data
        self.T = len(self.offsets)
        if self.T < 1:
            raise ValueError("threshold_range must span at least one threshold")
        self.plan = Plan()
        _lib.check(self.lib.chips_plan_init(C.byref(self.plan), geom.H, geom.W, self.k, self.h_bins,
                                            self.T, elem, geom.cx, geom.cy, geom.r), "chips_plan_init")
        # float32 caller arrays: numpy keeps float32 for the bin edges, bound / limit_0 and the
        # sigmoid of calculate_prob (np.histogram's bin_type = result_type(a)); everything else is
        # float64 arithmetic on the widened values
        self.f32_math = bool(f32_math) and elem == 4
        self.plan.f32_math = int(self.f32_math)
        # scratch tail behind the plan's workspace: outputs of the regional histograms
        # (`region_histogram`), so that they never overwrite the detection's own record
        self._tail = {}
        off = (int(self.plan.bytes) + 255) // 256 * 256
        for name, nbytes in (("minmax", 16), ("counts", self.h_bins * 8), ("density", self.h_bins * 8),
                             ("edges", (self.h_bins + 1) * 8), ("bound", self.h_bins * 8),
                             ("peaks", self.h_bins * 4), ("result", C.sizeof(Result))):
            self._tail[name] = off
            off = (off + nbytes + 255) // 256 * 256
        self.ws_bytes = off
        if ws is not None and ws.numel() >= off:
            self.ws = ws            # shared by the caller (detectors of one stream slot run in turn)
        else:
            with torch.cuda.device(self.device):
                self.ws = torch.empty(off, dtype=torch.uint8, device=self.device)
        self._pp = C.byref(self.plan)
        rect = (C.c_int32 * 4)()
        _lib.check(self.lib.chips_input_rect(self._pp, rect), "chips_input_rect")
        self.input_rect = tuple(int(v) for v in rect)     # x0, y0, x1, y1: all a detection reads of its input
        sp = (C.c_int32 * (4 * 4096))()
        ns = int(self.lib.chips_input_spans(self._pp, geom.cx, geom.cy, geom.r, sp, 4096))
        if ns < 1:
            raise _lib.ChipsNativeError(f"chips_input_spans failed with code {ns}")
        # (y0, y1, x0, x1) row intervals inside input_rect that hold everything a detection reads
        self.input_spans = [tuple(int(sp[4 * i + j]) for j in range(4)) for i in range(ns)]
        self.dtype = torch.float32 if elem == 4 else torch.float64
        self.graph_launches = 0        # kernels launched through CUDA-graph replays (not seen by the library counter)

    def upload(self, dst: torch.Tensor, src: torch.Tensor) -> int:
        """Host image -> device image, only `input_spans` (row intervals inside `input_rect`; the
        rest of `dst` is never read by a detection with this plan).  Enqueued on the current
        stream; returns the bytes copied."""
        x0, y0, x1, y1 = self.input_rect
        H, W = self.geom.H, self.geom.W
        if (src.dtype != dst.dtype or tuple(src.shape) != (H, W) or tuple(dst.shape) != (H, W)
                or not src.is_contiguous() or not dst.is_contiguous() or src.is_cuda or not dst.is_cuda):
            dst.copy_(src, non_blocking=True)
            return src.numel() * src.element_size()
        e = src.element_size()
        total = 0
        for (sy0, sy1, sx0, sx1) in self.input_spans:      # ~26 two-dimensional copies at 4096^2
            _lib.check(self.lib.chips_copy_rect(_ptr(dst), C.c_void_p(src.data_ptr()), W * e, sx0 * e, sy0,
                                                (sx1 - sx0) * e, sy1 - sy0, 1, _stream_ptr()), "chips_copy_rect")
            total += (sx1 - sx0) * e * (sy1 - sy0)
        return total

    # ------------------------------------------------------------------ workspace views
    def view(self, name: str, dtype, shape) -> torch.Tensor:
        off = int(getattr(self.plan, "off_" + name))
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.ws[off:off + n].view(dtype).reshape(shape)

    @property
    def filt(self):
        return self.view("filt", self.dtype, (self.geom.H, self.geom.W))

    @property
    def level(self):
        return self.view("level", torch.uint8, (self.geom.H, self.geom.W))

    @property
    def fill(self):
        return self.view("fill", torch.uint8, (self.geom.H, self.geom.W))

    @property
    def keep(self):
        return self.view("keep", torch.uint8, (self.geom.H, self.geom.W))

    @property
    def counts(self):
        return self.view("counts", torch.int64, (self.h_bins,))

    @property
    def density(self):
        return self.view("density", torch.float64, (self.h_bins,))

    @property
    def edges(self):
        return self.view("edges", torch.float64, (self.h_bins + 1,))

    @property
    def bound(self):
        return self.view("bound", torch.float64, (self.h_bins,))

    @property
    def peak_index(self):
        return self.view("peaks", torch.int32, (self.h_bins,))

    @property
    def probtab(self):
        return self.view("probtab", torch.int32, (self.T + 1, _lib.PROB_BINS))

    def result_bytes(self) -> torch.Tensor:
        off = int(self.plan.off_result)
        return self.ws[off:off + C.sizeof(Result)]

    def result(self) -> Result:
        """Device -> host read of the per-image record (synchronises the current stream)."""
        host = self.result_bytes().cpu().numpy().tobytes()
        return Result.from_buffer_copy(host)

    # ------------------------------------------------------------------ stages
    def _g(self):
        return self.geom.cx, self.geom.cy, self.geom.r

    def medfilt(self, image: torch.Tensor, out: torch.Tensor | None = None):
        out = self.filt if out is None else out
        _lib.check(self.lib.chips_medfilt2d(_ptr(image), _ptr(out), self._pp, *self._g(),
                                            _ptr(self.ws), _stream_ptr()), "chips_medfilt2d")
        if self.k == 1:
            _lib.check(self.lib.chips_minmax(_ptr(out), self._pp, *self._g(), _ptr(self.ws),
                                             _stream_ptr()), "chips_minmax")
        return out

    def histogram(self, filt: torch.Tensor | None = None, recompute_minmax: bool = False):
        filt = self.filt if filt is None else filt
        if recompute_minmax:
            _lib.check(self.lib.chips_minmax(_ptr(filt), self._pp, *self._g(), _ptr(self.ws),
                                             _stream_ptr()), "chips_minmax")
        _lib.check(self.lib.chips_histogram(_ptr(filt), self._pp, *self._g(), _ptr(self.ws),
                                            _stream_ptr()), "chips_histogram")

    def region_histogram(self, y0: int, y1: int, x0: int, x1: int, ht_peak_ratio: float, h_thresh):
        """`np.histogram(bins=h_bins, density=True)` + `find_peaks(height=h_thresh)` of the on-disk
        pixels of `filt[y0:y1, x0:x1]` (the body of the region loop of `Chips.extract_histograms`,
        chips.py:282-300): the same three kernels as the full-disk histogram, run on a plan whose
        ROI is the region and whose outputs live in the scratch tail.  Returns
        (density, bound, peak_indices, h_thresh) as host arrays, or None for an empty region."""
        p = self.plan
        rx0, ry0 = max(int(p.roi_x0), x0), max(int(p.roi_y0), y0)
        rx1, ry1 = min(int(p.roi_x0 + p.roi_w), x1), min(int(p.roi_y0 + p.roi_h), y1)
        if rx1 <= rx0 or ry1 <= ry0:
            return None
        q = Plan.from_buffer_copy(bytes(p))
        q.roi_x0, q.roi_y0, q.roi_w, q.roi_h = rx0, ry0, rx1 - rx0, ry1 - ry0
        for name, off in self._tail.items():
            setattr(q, "off_" + name, off)
        qp = C.byref(q)
        h_in = float("nan") if h_thresh is None else float(h_thresh)
        for fn, args in ((self.lib.chips_minmax, (_ptr(self.filt), qp, *self._g())),
                         (self.lib.chips_histogram, (_ptr(self.filt), qp, *self._g()))):
            _lib.check(fn(*args, _ptr(self.ws), _stream_ptr()), fn.__name__)
        _lib.check(self.lib.chips_select_threshold(qp, float(ht_peak_ratio), h_in, self._offp,
                                                   _ptr(self.ws), _stream_ptr()), "chips_select_threshold")

        def tail(name, dtype, count):
            o = self._tail[name]
            nb = count * torch.empty((), dtype=dtype).element_size()
            return self.ws[o:o + nb].view(dtype)

        res = Result.from_buffer_copy(tail("result", torch.uint8, C.sizeof(Result)).cpu().numpy().tobytes())
        if res.n_samples == 0:
            return None
        npk = int(res.n_peaks)
        return (tail("density", torch.float64, self.h_bins).cpu().numpy(),
                tail("bound", torch.float64, self.h_bins).cpu().numpy(),
                tail("peaks", torch.int32, self.h_bins)[:npk].cpu().numpy().astype(np.int64),
                np.float64(res.h_thresh))

    def window_histograms(self, windows, ht_peak_ratio: float, h_thresh):
        """Histogram + find_peaks of the on-disk pixels of `filt[r0:r1, c0:c1]` for every window
        (r0, r1, c0, c1) at once (`chips_window_histograms`: three launches, one warp per window for
        the selection).  Returns (density [n, h_bins], bound [n, h_bins], list of peak-index arrays,
        h_thresh) as host arrays; h_thresh = None lets the first window fix it (sticky)."""
        n = len(windows)
        rects = np.ascontiguousarray([[c0, r0, c1, r1] for (r0, r1, c0, c1) in windows], dtype=np.int32)
        hb = self.h_bins
        dev = self.device
        counts = torch.empty((n, hb), dtype=torch.int64, device=dev)
        density = torch.empty((n, hb), dtype=torch.float64, device=dev)
        bound = torch.empty((n, hb), dtype=torch.float64, device=dev)
        peaks = torch.empty((n, hb), dtype=torch.int32, device=dev)
        recs = torch.empty((n, C.sizeof(_lib.WindowRec)), dtype=torch.uint8, device=dev)
        ws = torch.empty(int(self.lib.chips_window_workspace_bytes(n)), dtype=torch.uint8, device=dev)
        h_in = float("nan") if h_thresh is None else float(h_thresh)
        _lib.check(self.lib.chips_window_histograms(
            _ptr(self.filt), self.geom.H, self.geom.W, self.elem, int(self.f32_math), *self._g(), n,
            rects.ctypes.data, hb,
            float(ht_peak_ratio), h_in, _ptr(counts), _ptr(density), _ptr(bound), _ptr(peaks), _ptr(recs),
            _ptr(ws), _stream_ptr()), "chips_window_histograms")
        rec_host = recs.cpu().numpy()
        rl = [_lib.WindowRec.from_buffer_copy(rec_host[i].tobytes()) for i in range(n)]
        pk = peaks.cpu().numpy()
        return (density.cpu().numpy(), bound.cpu().numpy(),
                [pk[i, :rl[i].n_peaks].astype(np.int64) for i in range(n)],
                np.float64(rl[0].h_thresh) if n else None, counts.cpu().numpy())

    def select_threshold(self, ht_peak_ratio: float, h_thresh):
        h_in = float("nan") if h_thresh is None else float(h_thresh)
        _lib.check(self.lib.chips_select_threshold(self._pp, float(ht_peak_ratio), h_in, self._offp,
                                                   _ptr(self.ws), _stream_ptr()),
                   "chips_select_threshold")

    def threshold_prob(self, porb_threshold: float, filt: torch.Tensor | None = None):
        filt = self.filt if filt is None else filt
        _lib.check(self.lib.chips_threshold_prob(_ptr(filt), self._pp, *self._g(),
                                                 float(porb_threshold), _ptr(self.ws), _stream_ptr()),
                   "chips_threshold_prob")

    def cleanup(self, area_threshold: float):
        if self.geom.r < 0:
            self.keep.copy_(self.level)      # synoptic: map = raw threshold mask
            return
        _lib.check(self.lib.chips_cleanup(self._pp, *self._g(), self.geom.disk_area,
                                          float(area_threshold), _ptr(self.ws), _stream_ptr()),
                   "chips_cleanup")

    def detect(self, image: torch.Tensor, ht_peak_ratio=5, h_thresh=None, porb_threshold=0.8,
               area_threshold=1e-3, maps: torch.Tensor | None = None):
        h_in = float("nan") if h_thresh is None else float(h_thresh)
        _lib.check(self.lib.chips_detect(_ptr(image), self._pp, *self._g(), float(ht_peak_ratio),
                                         h_in, self._offp, float(porb_threshold),
                                         float(area_threshold), _ptr(maps), _ptr(self.ws),
                                         _stream_ptr()), "chips_detect")

    def detect_graphed(self, image: torch.Tensor, ht_peak_ratio=5, h_thresh=None, porb_threshold=0.8,
                       area_threshold=1e-3, maps: torch.Tensor | None = None):
        """`detect` replayed from a CUDA graph.  One detection is ~200 small launches and the host
        needs ~1.2 ms to enqueue them, which caps a multi-stream series at ~800 images/s per
        process; the launch sequence is fixed for fixed buffers, so it is captured once per
        (input buffer, maps buffer, parameters) and replayed with a single graph launch.  The
        first call with a new key runs normally and captures for the following ones."""
        h_in = float("nan") if h_thresh is None else float(h_thresh)
        key = (image.data_ptr(), 0 if maps is None else maps.data_ptr(), float(ht_peak_ratio),
               None if h_thresh is None else h_in, float(porb_threshold), float(area_threshold))
        graphs = self.__dict__.setdefault("_graphs", {})
        entry = graphs.get(key)
        if entry is not None:
            entry[0].replay()
            self.graph_launches += entry[1]          # kernels inside the replayed graph
            return
        self.detect(image, ht_peak_ratio, h_thresh, porb_threshold, area_threshold, maps)
        cur = torch.cuda.current_stream()
        cap = self.__dict__.setdefault("_cap_stream", torch.cuda.Stream(device=self.device))
        cap.wait_stream(cur)
        g = torch.cuda.CUDAGraph()
        n0 = self.lib.chips_launch_count()
        # Python's cycle collector must not run inside the capture: finalising an unrelated CUDA
        # graph / tensor there (cudaGraphDestroy, cudaFree) is an operation that invalidates it.
        # Collect before, keep the collector off while capturing.
        import gc
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            # thread_local: CUDA calls of other threads (NCCL watchdog, samplers) do not break the capture
            with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                self.detect(image, ht_peak_ratio, h_thresh, porb_threshold, area_threshold, maps)
        finally:
            if gc_was_on:
                gc.enable()
        nodes = int(self.lib.chips_launch_count() - n0)   # counted while capturing, not executed
        cur.wait_stream(cap)
        if len(graphs) >= 16:                         # (buffer, parameters) keys seen so far: keep the 16 newest
            graphs.pop(next(iter(graphs)))
        graphs[key] = (g, nodes)
        self.graph_launches -= nodes                  # undo the capture's count in the library total

    # ------------------------------------------------------------------ outputs
    def expand_maps(self, which: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
        """uint8 [T,H,W]: which=0 final maps, 1 raw threshold masks, 2 hole-filled masks."""
        if out is None:
            out = torch.empty((self.T, self.geom.H, self.geom.W), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.chips_expand_maps(self._pp, which, _ptr(out), _ptr(self.ws),
                                              _stream_ptr()), "chips_expand_maps")
        return out

    def expand_map(self, t: int, which: int = 0, dtype=torch.float64) -> torch.Tensor:
        out = torch.empty((self.geom.H, self.geom.W), dtype=dtype, device=self.device)
        _lib.check(self.lib.chips_expand_map(self._pp, which, int(t), _ptr(out), out.element_size(),
                                             _ptr(self.ws), _stream_ptr()), "chips_expand_map")
        return out

    def pack_cube(self, which: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
        """float32 [H,W,T] region cube of `Chips.to_netcdf` (chips.py:570-586)."""
        if out is None:
            out = torch.empty((self.geom.H, self.geom.W, self.T), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.chips_pack_cube(self._pp, which, _ptr(out), _ptr(self.ws), _stream_ptr()),
                   "chips_pack_cube")
        return out

    def roots(self, t: int):
        roots = torch.empty((self.geom.H, self.geom.W), dtype=torch.int32, device=self.device)
        area2 = torch.empty((self.geom.H, self.geom.W), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.chips_roots(self._pp, *self._g(), int(t), _ptr(roots), _ptr(area2),
                                        _ptr(self.ws), _stream_ptr()), "chips_roots")
        return roots, area2

    def labels(self, t: int):
        """Canonical label image (raster order of first pixel) and 2*area per label (host)."""
        roots, area2 = self.roots(t)
        r = roots.cpu().numpy()
        a = area2.cpu().numpy()
        fg = r >= 0
        uniq, first, inv = np.unique(r[fg], return_index=True, return_inverse=True)
        lab = np.zeros(r.shape, dtype=np.int32)
        lab[fg] = inv.astype(np.int32) + 1
        return lab, a[fg][first].astype(np.int64)

    def pack_keep(self, entries: torch.Tensor, count: torch.Tensor) -> None:
        """Append (pixel index << 6 | keep) of every pixel with keep < T to `entries` (uint32/int32
        tensor on the device, or PINNED host memory: the kernel writes through the mapping, no
        device -> host copy of the 16.8 MB map) and leave their number in `count` (device)."""
        ptr = entries.data_ptr()
        _lib.check(self.lib.chips_pack_keep(self._pp, ptr, int(entries.numel()), _ptr(count), _ptr(self.ws),
                                            _stream_ptr()), "chips_pack_keep")

    def labels_periodic(self, t: int, periodic: bool = True):
        """8-connected components of the final map of threshold t (`keep <= t`) with the x axis
        periodic (synoptic maps: longitude wraps) - `chips_label_periodic`.  Returns (labels int32
        [H, W] numbered in raster order of the first pixel, pixel counts per label) on the host."""
        H, W = self.geom.H, self.geom.W
        dev = self.device
        lab = torch.empty((H, W), dtype=torch.int32, device=dev)
        cap = ((H + 1) // 2) * ((W + 1) // 2)
        areas = torch.empty(cap, dtype=torch.int32, device=dev)
        nlab = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(int(self.lib.chips_label_workspace_bytes(H, W)), dtype=torch.uint8, device=dev)
        _lib.check(self.lib.chips_label_periodic(_ptr(self.keep), H, W, 0, int(t), int(periodic), _ptr(lab),
                                                 _ptr(areas), cap, _ptr(nlab), _ptr(ws), _stream_ptr()),
                   "chips_label_periodic")
        n = int(nlab.item())
        return lab.cpu().numpy(), areas[:n].cpu().numpy().astype(np.int64)

    def contours(self, t: int):
        """Contours of the raw threshold mask `level <= t` exactly as the reference's
        cv2.findContours call sees it (chips.py:382-386): (contours, hierarchy [n, 4],
        twice_area [n]) in OpenCV's order, before the area cut."""
        from . import contours as _c
        if getattr(self, "_tracer", None) is None:
            self._tracer = ContourTracer(self.geom.W, self.geom.H, self.device)
        recs, pts = self._tracer.trace(self.level, 0, int(t))
        return _c.assemble(recs, pts)

    def contours_all(self):
        """`contours(t)` for every threshold of the detection, enqueued together and synchronised
        once: [(contours, hierarchy, twice_area)] for t = 0 .. T-1."""
        from . import contours as _c
        if getattr(self, "_tracer", None) is None:
            self._tracer = ContourTracer(self.geom.W, self.geom.H, self.device)
        return [_c.assemble(recs, pts) for recs, pts in self._tracer.trace_many(self.level, range(self.T), 0)]

    def prob_stack(self, lower_lim: float = 0.0, out: torch.Tensor | None = None,
                   accumulate: bool = False, dtype=torch.float64) -> torch.Tensor:
        if out is None:
            out = torch.zeros((self.geom.H, self.geom.W), dtype=dtype, device=self.device)
        _lib.check(self.lib.chips_prob_stack(self._pp, float(lower_lim), _ptr(out), out.element_size(),
                                             int(accumulate), _ptr(self.ws), _stream_ptr()),
                   "chips_prob_stack")
        return out


def to_device_image(data: np.ndarray, device, allow_demote: bool = True):
    """Host array -> (device tensor, elem).  float64 input whose values are all exactly
    float32-representable is stored as float32 (half the HBM traffic; every comparison the
    reference makes in float64 is done on the widened value, so results are identical)."""
    lib = _lib.load()
    arr = np.ascontiguousarray(data)
    if arr.dtype == np.float32:
        return torch.from_numpy(arr).to(device, non_blocking=False), 4
    if arr.dtype != np.float64:
        arr = arr.astype(np.float64)
    t64 = torch.from_numpy(arr).to(device)
    if not allow_demote:
        return t64, 8
    t32 = torch.empty(arr.shape, dtype=torch.float32, device=device)
    flag = torch.ones(1, dtype=torch.int32, device=device)
    _lib.check(lib.chips_demote_f64(_ptr(t64), _ptr(t32), t64.numel(), _ptr(flag), _stream_ptr()),
               "chips_demote_f64")
    if int(flag.item()) == 1:
        return t32, 4
    return t64, 8


class ContourTracer:
    """Device buffers of `chips_contours` (csrc/contours.cu) for one image size; capacities grow
    on demand (the call reports what it needed)."""

    def __init__(self, w: int, h: int, device=None, max_border_px: int = 1 << 18,
                 max_contours: int = 1 << 15, max_points: int = 1 << 19):
        from . import contours as _c
        self._c = _c
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.w, self.h = int(w), int(h)
        self.cap, self.max_contours, self.max_points = int(max_border_px), int(max_contours), int(max_points)
        self.start_rounds = 8
        self.calls = 0
        self._alloc()

    def _alloc(self):
        nbytes = int(self.lib.chips_contours_workspace_bytes(self.w, self.h, self.cap))
        if nbytes == 0:
            raise ValueError("bad contour geometry")
        with torch.cuda.device(self.device):
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.recs = torch.empty(self.max_contours * self._c.REC_DTYPE.itemsize, dtype=torch.uint8,
                                    device=self.device)
            self.points = torch.empty(2 * self.max_points, dtype=torch.int32, device=self.device)
            self.counts = torch.zeros(self._c.COUNTS_DTYPE.itemsize, dtype=torch.uint8, device=self.device)

    def trace(self, img: torch.Tensor, mode: int = 1, thr: int = 0):
        """img: uint8 [h, w] device tensor (rows contiguous).  Returns (records, points) on the host:
        records in order of discovery (`contours.REC_DTYPE`), points int32 [n_points, 2]."""
        if img.dtype != torch.uint8 or img.dim() != 2 or img.shape != (self.h, self.w) or img.stride(1) != 1:
            raise ValueError("contour input must be a uint8 [h, w] device image")
        if img.device != self.device:
            raise ValueError("contour input lives on another device")
        c = self._c
        rounds = self.start_rounds        # per call: pipeline masks settle in 2-3 rounds
        for _ in range(24):
            with torch.cuda.device(self.device):
                _lib.check(self.lib.chips_contours(
                    _ptr(img), self.w, self.h, int(img.stride(0)), int(mode), int(thr), self.cap,
                    rounds, _ptr(self.recs), self.max_contours, _ptr(self.points),
                    self.max_points, _ptr(self.counts), _ptr(self.ws), self.ws.numel(), _stream_ptr()),
                    "chips_contours")
                self.calls += 1
                cnt = self.counts.cpu().numpy().view(c.COUNTS_DTYPE)[0]
            if cnt["overflow"] & 1:
                self.cap = int(cnt["n_border_px"]) * 5 // 4 + 1024
            elif cnt["overflow"]:
                self.max_contours = max(self.max_contours, int(cnt["n_contours"]) * 5 // 4 + 64)
                self.max_points = max(self.max_points, int(cnt["n_points"]) * 5 // 4 + 64)
            elif cnt["changed"]:          # lattices of 1-pixel walls need O(image side) rounds
                if rounds >= 1 << 20:
                    break
                rounds *= 2
                continue
            else:
                n, m = int(cnt["n_contours"]), int(cnt["n_points"])
                recs = self.recs[:n * c.REC_DTYPE.itemsize].cpu().numpy().view(c.REC_DTYPE)
                pts = self.points[:2 * m].cpu().numpy().reshape(-1, 2)
                return recs, pts
            self._alloc()
        raise _lib.ChipsNativeError("chips_contours did not settle")

    def trace_many(self, img: torch.Tensor, thrs, mode: int = 0):
        """The masks `img <= t` for every t in `thrs` in one enqueue and ONE synchronisation (the
        passes of the T calls run back to back on the stream and share the workspace; each call has
        its own output buffers).  A mask whose call reports an overflow or unsettled starts is redone
        by `trace`.  Returns [(records, points)] in the order of `thrs`."""
        if img.dtype != torch.uint8 or img.dim() != 2 or img.shape != (self.h, self.w) or img.stride(1) != 1:
            raise ValueError("contour input must be a uint8 [h, w] device image")
        c = self._c
        thrs = [int(t) for t in thrs]
        T, rb = len(thrs), c.REC_DTYPE.itemsize
        if T == 0:
            return []
        with torch.cuda.device(self.device):
            recs = torch.empty((T, self.max_contours * rb), dtype=torch.uint8, device=self.device)
            points = torch.empty((T, 2 * self.max_points), dtype=torch.int32, device=self.device)
            counts = torch.zeros((T, c.COUNTS_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
            for i, t in enumerate(thrs):
                _lib.check(self.lib.chips_contours(
                    _ptr(img), self.w, self.h, int(img.stride(0)), int(mode), t, self.cap,
                    self.start_rounds, _ptr(recs[i]), self.max_contours, _ptr(points[i]),
                    self.max_points, _ptr(counts[i]), _ptr(self.ws), self.ws.numel(), _stream_ptr()),
                    "chips_contours")
                self.calls += 1
            cnt = counts.cpu().numpy().view(c.COUNTS_DTYPE).reshape(T)          # the one synchronisation
            good = (cnt["overflow"] == 0) & (cnt["changed"] == 0)
            n_max = int(cnt["n_contours"][good].max()) if good.any() else 0
            m_max = int(cnt["n_points"][good].max()) if good.any() else 0
            recs_h = recs[:, :n_max * rb].cpu().numpy()
            pts_h = points[:, :2 * m_max].cpu().numpy()
        out = []
        for i, t in enumerate(thrs):
            if not good[i]:
                out.append(self.trace(img, mode, t))
                continue
            n, m = int(cnt["n_contours"][i]), int(cnt["n_points"][i])
            out.append((np.ascontiguousarray(recs_h[i, :n * rb]).view(c.REC_DTYPE),
                        np.ascontiguousarray(pts_h[i, :2 * m]).reshape(-1, 2)))
        return out


def find_contours(mask: np.ndarray, device=None):
    """`cv2.findContours(mask, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)` on the GPU: (contours,
    hierarchy [1, n, 4] or None)."""
    from . import contours as _c
    device = require_cuda(device)
    m = np.ascontiguousarray(mask)
    if m.ndim != 2 or m.dtype != np.uint8:
        raise ValueError("mask must be a 2-D uint8 array")
    with torch.cuda.device(device):
        img = torch.from_numpy(m).to(device)
        tracer = ContourTracer(m.shape[1], m.shape[0], device)
        recs, pts = tracer.trace(img, 1, 0)
    cs, hier, _ = _c.assemble(recs, pts)
    return tuple(cs), (hier[None] if len(cs) else None)


def row_cosine(map_: np.ndarray, map0: np.ndarray, device=None):
    """Per-row cosine similarities and their nanmean (`chips_row_cosine`).  Returns
    (rows [H] float64 with NaN for skipped rows, mean, rows_used)."""
    device = require_cuda(device)
    a = np.ascontiguousarray(map_)
    b = np.ascontiguousarray(map0)
    if a.ndim != 2 or a.shape != b.shape:
        raise ValueError("map and map0 must be 2-D arrays of the same shape")
    dt = np.float32 if (a.dtype == np.float32 and b.dtype == np.float32) else np.float64
    with torch.cuda.device(device):
        ta = torch.from_numpy(a.astype(dt, copy=False)).to(device)
        tb = torch.from_numpy(b.astype(dt, copy=False)).to(device)
        out = torch.empty(a.shape[0] + 2, dtype=torch.float64, device=device)
        _lib.check(_lib.load().chips_row_cosine(_ptr(ta), _ptr(tb), a.shape[0], a.shape[1],
                                                ta.element_size(), _ptr(out), _stream_ptr()),
                   "chips_row_cosine")
        h = out.cpu().numpy()
    return h[:-2], h[-2], int(h[-1])


def medfilt2d(data: np.ndarray, k: int, device=None) -> np.ndarray:
    """`scipy.signal.medfilt2d(data, k)` on the GPU (zero padded, exact)."""
    device = require_cuda(device)
    img, elem = to_device_image(data, device, allow_demote=False)
    H, W = data.shape
    det = Detector(Geometry(H, W, 0, 0, -1), k, 1, 0, 1, elem, device)
    out = torch.empty_like(img)
    det.medfilt(img, out)
    return out.cpu().numpy()
