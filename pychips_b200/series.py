"""Time series / batch detection: one full-disk image per detection, images sharded
round-robin across ranks (one process per GPU).  Each rank keeps `n_streams` detections in
flight on their own compute streams; host->device copies run one image ahead on a copy stream
(two input buffers per slot) and device->host copies read a device-side snapshot of the outputs
on another, so neither direction of PCIe traffic stalls a compute stream.

The reference has no batch API (one `Chips` per `RegisterAIA`, `/root/reference/chips/chips.py:
110-150`); this is the multi-image driver SURVEY.md §8(e) asks for.  The only exchange step is
at the end of a batch: `all_gather` of the per-image statistics and one `all_reduce(sum)` of the
per-rank accumulated probability map (plots.py:305-322 arithmetic) for the uncertainty estimate.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, engine
from .engine import Detector, Geometry

STAT_FIELDS = 4          # limit_0, h_thresh, n_peaks, median_errors  (+ 2*T lims/probs, + T kept)


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Round-robin partition of a batch (image i -> rank i % world)."""
    return list(range(rank, n_items, world))


@dataclass
class SeriesResult:
    indices: list
    limit_0: np.ndarray            # [n]
    h_thresh: np.ndarray           # [n]
    n_peaks: np.ndarray            # [n]
    limits: np.ndarray             # [n, T]
    probs: np.ndarray              # [n, T]
    n_kept: np.ndarray             # [n, T]
    keep: list = field(default_factory=list)   # per image uint8 [H,W] level-encoded maps (host; decoded
                                               # on first access when the runner downloads them sparse)
    prob_map_sum: torch.Tensor | None = None   # device fp32 [H,W], summed over the batch


def decode_keep(entries: np.ndarray, H: int, W: int, T: int) -> np.ndarray:
    """Dense level-encoded map from the words of `chips_pack_keep` ((pixel index << 6) | keep)."""
    out = np.full(H * W, T, dtype=np.uint8)
    e = np.asarray(entries).view(np.uint32)
    out[e >> 6] = (e & 63).astype(np.uint8)
    return out.reshape(H, W)


class SparseKeepList:
    """`SeriesResult.keep` of a runner with `sparse_keep=True`: list-like, image i is decoded from
    its packed words on first access (`packed(i)` returns the words themselves)."""

    def __init__(self, entries_host, counts_host, H, W, T):
        self._e, self._c, self._shape, self._T = entries_host, counts_host, (H, W), T
        self._cache = {}

    def __len__(self):
        return len(self._e)

    def is_dense(self, i) -> bool:
        """More pixels in some map than the buffer has words: it holds the dense map instead."""
        return int(self._c[i]) > self._e[i].numel()

    def packed(self, i) -> np.ndarray:
        """The words of image i (or, for `is_dense(i)`, the dense map as H*W bytes)."""
        if self.is_dense(i):
            H, W = self._shape
            return self._e[i].numpy().view(np.uint8)[:H * W]
        return self._e[i].numpy()[:int(self._c[i])].view(np.uint32)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if i not in self._cache:
            self._cache[i] = (self.packed(i).reshape(self._shape).copy() if self.is_dense(i)
                              else decode_keep(self.packed(i), *self._shape, self._T))
        return self._cache[i]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class SeriesRunner:
    """Runs `chips_detect` over a list of host images on ONE GPU with `n_streams` in flight."""

    def __init__(self, H: int, W: int, medfilt_kernel: int = 51, h_bins: int = 500,
                 threshold_range=(0, 20), ht_peak_ratio=5, porb_threshold=0.8,
                 area_threshold=1e-3, masked: bool = True, device=None, n_streams: int = 2,
                 elem: int = 4, want_keep: bool = True, want_prob_map: bool = True,
                 want_maps: bool = False, sparse_keep: bool = True, f32_math: bool = False):
        self.device = engine.require_cuda(device)
        self.H, self.W = int(H), int(W)
        self.k, self.h_bins = int(medfilt_kernel), int(h_bins)
        self.t0, self.t1 = threshold_range[0], threshold_range[1]
        self.T = len(np.arange(self.t0, self.t1))
        self.ratio, self.porb, self.area_thr = ht_peak_ratio, porb_threshold, area_threshold
        self.masked = masked
        self.elem = elem
        # float32 images are a storage format here: the arithmetic is float64 on the widened values
        # (what the reference does with float64 arrays of the same values).  f32_math=True gives
        # what the reference does when it is handed float32 arrays (numpy's float32 histogram)
        self.f32_math = bool(f32_math) and elem == 4
        self.want_keep, self.want_prob_map, self.want_maps = want_keep, want_prob_map, want_maps
        # the result map travels as one word per pixel that is in some map (~3 % of a disk image)
        # instead of H*W bytes; needs the word format of chips_pack_keep
        self.sparse_keep = bool(sparse_keep and want_keep and self.H * self.W <= (1 << 26)
                                and self.T <= 64 and self.W % 4 == 0)
        # H*W/4 words: if more pixels than that are in some map the buffer takes the dense map
        self.sparse_cap = (self.H * self.W + 3) // 4
        self.dtype = torch.float32 if elem == 4 else torch.float64
        self.n_streams = n_streams
        self._slots = {}
        self._slot_ws = {}
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
            self.dev_in = [torch.empty((self.H, self.W), dtype=self.dtype, device=self.device)
                           for _ in range(n_streams)]
            self.maps = [torch.empty((self.T, self.H, self.W), dtype=torch.uint8, device=self.device)
                         if want_maps else None for _ in range(n_streams)]
            # one summed probability map per host buffer set: a batch's map is zeroed when the batch
            # is submitted and belongs to that batch alone (batches overlap by one in flight)
            self._prob_sums = [torch.zeros((self.H, self.W), dtype=torch.float32, device=self.device)
                               for _ in range(self.n_host_sets if want_prob_map else 1)]
            self.prob_sum = self._prob_sums[0]           # the map of the most recently submitted batch

    def _detector(self, slot: int, r: int) -> Detector:
        key = (slot, r)
        det = self._slots.get(key)
        if det is None:
            # one plan (and one set of captured graphs) per (stream slot, radius), but ONE workspace
            # per slot, sized for a disk that covers the whole image: a series with a jittered
            # radius does not multiply the 0.5 GB workspaces
            c = int(self.H / 2)
            ws = self._slot_ws.get(slot)
            if ws is None and self.masked:
                big = Detector(Geometry(self.H, self.W, c, c, self.H + self.W), self.k, self.h_bins,
                               self.t0, self.t1, self.elem, self.device)
                ws = self._slot_ws[slot] = big.ws
            g = Geometry(self.H, self.W, c, c, int(r) if self.masked else -1)
            det = Detector(g, self.k, self.h_bins, self.t0, self.t1, self.elem, self.device, ws=ws,
                           f32_math=self.f32_math)
            self._slot_ws.setdefault(slot, det.ws)
            self._slots[key] = det
        return det

    def run(self, images, radii, indices=None, h_thresh=None) -> SeriesResult:
        """images: list of host arrays (pinned torch tensors or numpy, dtype matching `elem`);
        radii: per-image pixel_radius.  Copies each image to the device, runs the fused
        detection, copies the record (and the level-encoded map) back."""
        return self.submit(images, radii, indices, h_thresh).result()

    def submit(self, images, radii, indices=None, h_thresh=None) -> "PendingSeries":
        """Enqueue a batch without waiting for it: the returned handle's `result()` blocks until the
        batch's records and maps are on the host.  Batches submitted back to back keep all streams
        busy (no pipeline drain between them); at most `n_host_sets` (2) may be in flight, each
        with its own pinned result buffers.  The `keep` arrays of a result are VIEWS of those
        pinned buffers: they stay valid until the second following `submit` of the same batch
        size — copy them if they must live longer."""
        n = len(images)
        indices = list(range(n)) if indices is None else list(indices)
        rec_bytes = C.sizeof(_lib.Result)
        sets = self.__dict__.setdefault("_host_sets", {})
        ring = sets.get(n)
        if ring is None:                                # pinned result buffers, reused round-robin
            def keep_bufs():
                if not self.want_keep:
                    return []
                if self.sparse_keep:        # packed words, written by the kernel through the mapping
                    return [torch.empty(self.sparse_cap, dtype=torch.int32).pin_memory() for _ in range(n)]
                return [torch.empty((self.H, self.W), dtype=torch.uint8).pin_memory() for _ in range(n)]
            ring = sets[n] = dict(next=0, bufs=[
                (torch.empty((n, rec_bytes), dtype=torch.uint8).pin_memory(), keep_bufs(), [None],
                 torch.zeros(n, dtype=torch.int32).pin_memory())
                for _ in range(self.n_host_sets)])
        rec_host, keep_host, owner, cnt_host = ring["bufs"][ring["next"]]
        ring["next"] = (ring["next"] + 1) % self.n_host_sets
        if owner[0] is not None and not owner[0]._done:
            owner[0].result()                            # the set is still in flight: finish it first
        self._prob_turn = (getattr(self, "_prob_turn", -1) + 1) % len(self._prob_sums)
        prob_sum = self.prob_sum = self._prob_sums[self._prob_turn]
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            self._ensure_copy_pipeline()
            for s in self.streams + [self._h2d, self._d2h, self._acc]:
                s.wait_stream(cur)
            if self.want_prob_map:
                # the batch's own map starts at zero (on the accumulation stream, i.e. after every
                # accumulation of the batch that used this buffer two submits ago)
                with torch.cuda.stream(self._acc):
                    prob_sum.zero_()
            for i in range(n):
                slot = i % self.n_streams
                st = self.streams[slot]
                det = self._detector(slot, radii[i])
                src = images[i]
                if isinstance(src, np.ndarray):
                    src = torch.from_numpy(src)
                par = self._parity[slot]
                self._parity[slot] ^= 1
                din = self._dev_in2[slot][par]
                # host -> device on the copy stream, one image ahead of the slot's compute stream
                # (two input buffers per slot; a buffer is reused once its detection has read it)
                with torch.cuda.stream(self._h2d):
                    self._h2d.wait_event(self._in_free[slot][par])
                    if self.roi_copy:      # only the rectangle the kernels read (62 % at 4096^2 / r = 1577)
                        self.h2d_bytes += det.upload(din, src)
                    else:
                        din.copy_(src, non_blocking=True)
                        self.h2d_bytes += src.numel() * src.element_size()
                    self._in_ready[slot][par].record(self._h2d)
                with torch.cuda.stream(st):
                    st.wait_event(self._in_ready[slot][par])
                    st.wait_event(self._acc_done[slot])      # the slot's previous map has been summed
                    # (a CUDA-graph replay: ~200 launches cost the host 1.2 ms otherwise)
                    det.detect_graphed(din, self.ratio, h_thresh, self.porb, self.area_thr, self.maps[slot])
                    self._in_free[slot][par].record(st)
                if self.want_prob_map:
                    # the batch probability map is summed on ONE stream in image order: no two
                    # accumulations overlap (race-free) and the sum is reproducible
                    with torch.cuda.stream(self._acc):
                        self._acc.wait_event(self._in_free[slot][par])
                        det.prob_stack(0.0, prob_sum, accumulate=True)
                        self._acc_done[slot].record(self._acc)
                with torch.cuda.stream(st):
                    # snapshot of the outputs (device -> device), so that the slot's next detection
                    # does not wait for the device -> host copy
                    st.wait_event(self._stage_free[slot])
                    self._stage_rec[slot].copy_(det.result_bytes(), non_blocking=True)
                    if self.sparse_keep:      # packed on the device; the copy stream moves the words
                        det.pack_keep(self._stage_sparse[slot], self._stage_cnt[slot])
                    elif self.want_keep:
                        self._stage_keep[slot].copy_(det.keep, non_blocking=True)
                    self._snap[slot].record(st)
                with torch.cuda.stream(self._d2h):
                    self._d2h.wait_event(self._snap[slot])
                    rec_host[i].copy_(self._stage_rec[slot], non_blocking=True)
                    if self.sparse_keep:
                        # a copy whose size only the device knows: 8 CTAs store the words (and their
                        # number) into the pinned buffers through the mapping
                        _lib.check(self._lib.chips_copy_words(
                            self._stage_sparse[slot].data_ptr(), keep_host[i].data_ptr(),
                            self._stage_cnt[slot].data_ptr(), self.sparse_cap,
                            cnt_host[i:i + 1].data_ptr(), engine._stream_ptr()), "chips_copy_words")
                    elif self.want_keep:
                        keep_host[i].copy_(self._stage_keep[slot], non_blocking=True)
                    self._stage_free[slot].record(self._d2h)
            for s in self.streams + [self._h2d, self._d2h, self._acc]:
                cur.wait_stream(s)
            done = torch.cuda.Event()
            done.record(cur)
        handle = PendingSeries(self, indices, rec_host, keep_host, done,
                               prob_sum if self.want_prob_map else None,
                               cnt_host if self.sparse_keep else None)
        owner[0] = handle
        return handle

    n_host_sets = 2
    roi_copy = True        # upload only Detector.input_rect of every image
    h2d_bytes = 0          # bytes uploaded so far (bench.py reports them per step)

    def _ensure_copy_pipeline(self):
        """Copy streams, double input buffers and output snapshots (created on first use)."""
        if getattr(self, "_h2d", None) is not None:
            return
        ns, dev = self.n_streams, self.device
        self._h2d = torch.cuda.Stream(device=dev)
        self._d2h = torch.cuda.Stream(device=dev)
        self._acc = torch.cuda.Stream(device=dev)
        self._acc_done = [torch.cuda.Event() for _ in range(ns)]
        self._dev_in2 = [[self.dev_in[s], torch.empty_like(self.dev_in[s])] for s in range(ns)]
        self._parity = [0] * ns
        ev = lambda: torch.cuda.Event()
        self._in_free = [[ev(), ev()] for _ in range(ns)]
        self._in_ready = [[ev(), ev()] for _ in range(ns)]
        self._stage_free = [ev() for _ in range(ns)]
        self._snap = [ev() for _ in range(ns)]
        self._stage_rec = [torch.empty(C.sizeof(_lib.Result), dtype=torch.uint8, device=dev) for _ in range(ns)]
        self._stage_keep = [torch.empty((self.H, self.W), dtype=torch.uint8, device=dev)
                            if self.want_keep and not self.sparse_keep else None for _ in range(ns)]
        self._stage_cnt = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(ns)]
        self._stage_sparse = [torch.empty(self.sparse_cap, dtype=torch.int32, device=dev)
                              if self.sparse_keep else None for _ in range(ns)]
        self._lib = _lib.load()
        cur = torch.cuda.current_stream()
        for s in range(ns):                      # events start "signalled"
            for p in range(2):
                self._in_free[s][p].record(cur)
            self._stage_free[s].record(cur)
            self._acc_done[s].record(cur)


class PendingSeries:
    """A submitted batch (`SeriesRunner.submit`)."""

    def __init__(self, runner, indices, rec_host, keep_host, done, prob_sum=None, cnt_host=None):
        # no strong reference to the runner: the runner keeps this handle (its pinned buffers are
        # still in flight), a reference back would make a cycle that only the cycle collector frees
        self._T, self._indices = runner.T, indices
        self._rec_host, self._keep_host, self._event = rec_host, keep_host, done
        self._prob_sum = prob_sum
        self._cnt_host, self._shape = cnt_host, (runner.H, runner.W)
        self._done, self._result = False, None

    def result(self) -> SeriesResult:
        if self._done:
            return self._result
        self._event.synchronize()
        n, T = len(self._indices), self._T
        recs = [_lib.Result.from_buffer_copy(self._rec_host[i].numpy().tobytes()) for i in range(n)]
        for rc in recs:
            _lib.check_record(rc)
        self._result = SeriesResult(
            indices=self._indices,
            limit_0=np.array([rc.limit_0 for rc in recs]),
            h_thresh=np.array([rc.h_thresh for rc in recs]),
            n_peaks=np.array([rc.n_peaks for rc in recs]),
            limits=np.array([[rc.limits[t] for t in range(T)] for rc in recs]).reshape(n, T),
            probs=np.array([[rc.probs[t] for t in range(T)] for rc in recs]).reshape(n, T),
            n_kept=np.array([[rc.n_kept[t] for t in range(T)] for rc in recs]).reshape(n, T),
            keep=(SparseKeepList(self._keep_host, self._cnt_host.numpy(), *self._shape, T)
                  if self._cnt_host is not None else [k.numpy() for k in self._keep_host]),
            prob_map_sum=self._prob_sum,
        )
        self._done = True
        return self._result


def pack_stats(res: SeriesResult, T: int) -> torch.Tensor:
    """[n, 3 + 3T] float64: image index, limit_0, n_peaks, limits, probs, n_kept."""
    n = len(res.indices)
    out = np.zeros((n, 3 + 3 * T), dtype=np.float64)
    if n:
        out[:, 0] = res.indices
        out[:, 1] = res.limit_0
        out[:, 2] = res.n_peaks
        out[:, 3:3 + T] = res.limits
        out[:, 3 + T:3 + 2 * T] = res.probs
        out[:, 3 + 2 * T:] = res.n_kept
    return torch.from_numpy(out)


def gather_series(stats: torch.Tensor, prob_sum: torch.Tensor | None, n_total: int, T: int,
                  group=None):
    """The one exchange step of a batch: all_gather of the per-image statistics (padded to the
    largest shard) and all_reduce(sum) of the probability map.  Works on NCCL (device tensors)
    and on gloo (CPU tensors; used by the world_size-2 CPU tests).  `prob_sum` (this rank's map)
    is NOT modified: the reduction runs on a copy, which is returned — reducing in place would
    re-sum an already reduced map (x world) whenever the caller's buffer lives on."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    dev = prob_sum.device if prob_sum is not None else stats.device
    if dist.get_backend(group) == "nccl":
        stats = stats.to(dev)
    pad = torch.full((per, 3 + 3 * T), float("nan"), dtype=torch.float64, device=stats.device)
    pad[:stats.shape[0]] = stats
    bucket = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bucket, pad, group=group)
    total = None
    if prob_sum is not None:
        total = prob_sum.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    allstats = torch.cat(bucket).cpu().numpy()
    allstats = allstats[~np.isnan(allstats[:, 0])]
    order = np.argsort(allstats[:, 0], kind="stable")
    return allstats[order], total
