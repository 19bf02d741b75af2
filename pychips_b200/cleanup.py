"""Drop-in for `chips.cleanup.CleanFl` (`/root/reference/chips/cleanup.py:37-200`): CHIMERA-style
coronal-hole candidate bitmaps from three co-registered wavelengths, computed by the sm_100a
kernels in `csrc/fl.cu` through the C ABI (`chips_fl_candidates`).  No CPU fallback.

Scope: the three maps must already have the requested resolution (the reference's
`skimage.transform.resize` is the identity then); `produce_summary_plots` is presentation and is
not provided.
"""
from __future__ import annotations

import ctypes as C
from argparse import Namespace
from typing import Tuple

import numpy as np
import torch

from . import _lib, engine

DEFAULT_PARAMS = dict(                      # cleanup.py:52-60
    clip_negative=0,
    clip_211_log=[0.8, 2.7],
    clip_193_log=[1.4, 3.0],
    clip_171_log=[1.2, 3.9],
    bmmix_value=0.6357,
    bmhot_value=0.7,
    bmcool_value=1.5102,
)


class CleanFl(object):
    """Same constructor, attributes and return values as the reference class."""

    def __init__(self, aia, resolution: int = 4096, params: dict = DEFAULT_PARAMS, device=None):
        self.aia = aia
        self.base_folder = getattr(aia, "save_dir", None)
        self.params = Namespace(**params)
        self.resolution = resolution
        self.device = engine.require_cuda(device)
        self.bits = None                 # device uint8 [H][W]: OR of _lib.FL_* per pixel
        self.stats = None                # _lib.FlStats (host copy)

    def _c_params(self) -> _lib.FlParams:
        p = self.params
        two = C.c_double * 2
        return _lib.FlParams(float(p.clip_negative), two(*map(float, p.clip_211_log)),
                             two(*map(float, p.clip_193_log)), two(*map(float, p.clip_171_log)),
                             float(p.bmmix_value), float(p.bmhot_value), float(p.bmcool_value))

    def _data(self, smap, wavelength: int, resolution: int) -> np.ndarray:
        smap = smap if smap else self.aia.datasets[wavelength][resolution].normalized
        data = np.asarray(smap.data)
        if data.shape != (resolution, resolution):
            raise NotImplementedError(
                f"{wavelength} A map is {data.shape}, not {(resolution, resolution)}: the resampling "
                "step (skimage.transform.resize, cleanup.py:113-130) is outside the accelerated path")
        return data

    def create_coronal_hole_candidates(self, smap193=None, smap171=None, smap211=None,
                                       disk_mask: np.ndarray = None, resolution: int = None) -> Tuple:
        """cleanup.py:69-200.  Returns (candidate_bitmap, bmmix, bmhot, bmcool, disk_mask); the
        first four are the UNMASKED float32 bitmaps, the attributes hold the masked copies."""
        resolution = resolution if resolution else self.resolution
        imgs = [self._data(smap171, 171, resolution), self._data(smap193, 193, resolution),
                self._data(smap211, 211, resolution)]
        f32 = all(a.dtype == np.float32 for a in imgs)
        npdt = np.float32 if f32 else np.float64
        lib = _lib.load()
        radius = int(self.aia.datasets[193][self.resolution].pixel_radius)
        c = int(resolution / 2)
        with torch.cuda.device(self.device):
            dev = [torch.from_numpy(np.ascontiguousarray(a, dtype=npdt)).to(self.device) for a in imgs]
            bits = torch.empty((resolution, resolution), dtype=torch.uint8, device=self.device)
            stats = torch.zeros(C.sizeof(_lib.FlStats), dtype=torch.uint8, device=self.device)
            ws = torch.empty(max(8, int(lib.chips_fl_workspace_bytes(resolution, resolution))),
                             dtype=torch.uint8, device=self.device)
            prm = self._c_params()
            _lib.check(lib.chips_fl_candidates(
                engine._ptr(dev[0]), engine._ptr(dev[1]), engine._ptr(dev[2]), resolution, resolution,
                dev[0].element_size(), c, c, radius, C.byref(prm), engine._ptr(bits), engine._ptr(stats),
                engine._ptr(ws), engine._stream_ptr()), "chips_fl_candidates")
            host = bits.cpu().numpy()
            self.stats = _lib.FlStats.from_buffer_copy(stats.cpu().numpy().tobytes())
        self.bits = bits

        def plane(flag):
            return ((host & flag) != 0).astype(np.float32)

        bmmix, bmhot, bmcool = plane(_lib.FL_MIX), plane(_lib.FL_HOT), plane(_lib.FL_COOL)
        candidate_bitmap = plane(_lib.FL_CAND)
        user_mask = disk_mask is not None
        if not user_mask:
            disk_mask = plane(_lib.FL_DISK).astype(np.float64)          # cleanup.py:101-111
        self.candidate_bitmap, self.bmmix, self.bmhot, self.bmcool = (
            candidate_bitmap * disk_mask, bmmix * disk_mask, bmhot * disk_mask, bmcool * disk_mask)
        # device bit read by the detection hook (chips.py:376-377, then `.astype(np.uint8)` at :383):
        # the STORED candidate bitmap, i.e. candidate AND mask
        cand = torch.tensor(_lib.FL_CAND, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            if user_mask:
                on = torch.from_numpy(np.ascontiguousarray(
                    self.candidate_bitmap.astype(np.uint8) != 0)).to(self.device)
            else:
                on = (bits & _lib.FL_DISK) != 0
            self.bits = torch.where(on, bits, bits & ~cand)
        return candidate_bitmap, bmmix, bmhot, bmcool, disk_mask

    def produce_summary_plots(self, *a, **k) -> None:
        raise NotImplementedError("plotting is outside the accelerated path (SURVEY.md §2 row 6)")
