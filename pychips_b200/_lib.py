"""ctypes binding of the C-ABI library (`include/chips_b200.h`).

There is deliberately NO fallback: if `libchips_b200.so` is missing the import of the
product path fails loudly (build it with `python -m pychips_b200.build`)."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB

MAX_T = 64
PROB_BINS = 20


class Result(C.Structure):
    """Mirror of `chips_result`."""
    _fields_ = [
        ("first_edge", C.c_double), ("last_edge", C.c_double), ("h_thresh", C.c_double),
        ("limit_0", C.c_double), ("limits", C.c_double * MAX_T), ("probs", C.c_double * MAX_T),
        ("n_samples", C.c_int64), ("n_peaks", C.c_int32), ("n_thresholds", C.c_int32),
        ("n_kept", C.c_int32 * MAX_T), ("n_comp", C.c_int32 * MAX_T),
        ("median_errors", C.c_int32), ("n_nonfinite", C.c_int32),
    ]


class Plan(C.Structure):
    """Mirror of `chips_plan`."""
    _fields_ = [(n, C.c_int32) for n in
                ("H", "W", "k", "h_bins", "T", "elem", "masked", "f32_math",
                 "roi_x0", "roi_y0", "roi_w", "roi_h", "bp_pitch", "bp_rows",
                 "blk_rows", "blk_nbx", "blk_nby", "blk_bucket")] + \
               [(n, C.c_size_t) for n in
                ("off_order", "off_bucket", "off_bstar", "off_rres", "off_filt", "off_minmax",
                 "off_counts", "off_density", "off_edges", "off_bound", "off_peaks",
                 "off_result", "off_level", "off_fill", "off_keep", "off_probtab",
                 "off_parent_b", "off_parent_c", "off_area", "off_pending", "off_list2", "off_lvl",
                 "off_tiles", "off_ovf", "bytes")]


class WindowRec(C.Structure):
    _fields_ = [("first_edge", C.c_double), ("last_edge", C.c_double), ("h_thresh", C.c_double),
                ("n_samples", C.c_int64), ("n_peaks", C.c_int32), ("pad_", C.c_int32)]


def check_record(res) -> None:
    """Raise for a detection record that must not be used (median self-check, non-finite input)."""
    if res.n_nonfinite:
        raise ChipsNativeError(
            f"the image holds NaN/Inf inside the region the detection reads ({res.n_nonfinite} hits): "
            "scipy.signal.medfilt2d is undefined there (chips.py:214-216); fill or mask the data first")
    if res.median_errors:
        raise ChipsNativeError(f"median filter self-check failed for {res.median_errors} pixels")


class FlParams(C.Structure):
    """Mirror of `chips_fl_params` (defaults: /root/reference/chips/cleanup.py:52-60)."""
    _fields_ = [("clip_negative", C.c_double), ("clip_211_log", C.c_double * 2),
                ("clip_193_log", C.c_double * 2), ("clip_171_log", C.c_double * 2),
                ("bmmix_value", C.c_double), ("bmhot_value", C.c_double), ("bmcool_value", C.c_double)]


class FlStats(C.Structure):
    """Mirror of `chips_fl_stats`."""
    _fields_ = [(n, C.c_double) for n in ("mean171", "mean193", "mean211", "thr_mix", "thr_hot", "thr_cool")] + \
               [(n, C.c_int64) for n in ("n_mix", "n_hot", "n_cool", "n_cand", "n_near")]


FL_MIX, FL_HOT, FL_COOL, FL_CAND, FL_DISK = 1, 2, 4, 8, 16

_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
_PP = C.POINTER(Plan)

SIGNATURES = {
    "chips_version": ([], C.c_int),
    "chips_launch_count": ([], C.c_longlong),
    "chips_medfilt2d_stages": ([_vp, _vp, _PP, _i, _i, _i, _vp, _vp, _i], C.c_int),
    "chips_plan_init": ([_PP, _i, _i, _i, _i, _i, _i, _i, _i, _i], C.c_int),
    "chips_disk_mask": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "chips_medfilt2d": ([_vp, _vp, _PP, _i, _i, _i, _vp, _vp], C.c_int),
    "chips_medfilt2d_bruteforce": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "chips_minmax": ([_vp, _PP, _i, _i, _i, _vp, _vp], C.c_int),
    "chips_histogram": ([_vp, _PP, _i, _i, _i, _vp, _vp], C.c_int),
    "chips_select_threshold": ([_PP, _d, _d, _vp, _vp, _vp], C.c_int),
    "chips_threshold_prob": ([_vp, _PP, _i, _i, _i, _d, _vp, _vp], C.c_int),
    "chips_cleanup": ([_PP, _i, _i, _i, _d, _d, _vp, _vp], C.c_int),
    "chips_expand_maps": ([_PP, _i, _vp, _vp, _vp], C.c_int),
    "chips_expand_map": ([_PP, _i, _i, _vp, _i, _vp, _vp], C.c_int),
    "chips_roots": ([_PP, _i, _i, _i, _i, _vp, _vp, _vp, _vp], C.c_int),
    "chips_prob_stack": ([_PP, _d, _vp, _i, _i, _vp, _vp], C.c_int),
    "chips_detect": ([_vp, _PP, _i, _i, _i, _d, _d, _vp, _d, _d, _vp, _vp, _vp], C.c_int),
    "chips_window_workspace_bytes": ([_i], C.c_size_t),
    "chips_window_histograms": ([_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _d, _d, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp], C.c_int),
    "chips_label_workspace_bytes": ([_i, _i], C.c_size_t),
    "chips_label_periodic": ([_vp, _i, _i, _i, _i, _i, _vp, _vp, C.c_longlong, _vp, _vp, _vp], C.c_int),
    "chips_pack_keep": ([_PP, _vp, C.c_uint32, _vp, _vp, _vp], C.c_int),
    "chips_copy_words": ([_vp, _vp, _vp, C.c_uint32, _vp, _vp], C.c_int),
    "chips_batch_create": ([_vp, _i, _i, _i, _i, _i, _i, _i, _i], C.c_int),
    "chips_detect_batch": ([_vp, _i, _vp, _i, _vp, _vp, _vp, _d, _d, _vp, _d, _d, _vp, _vp, _vp, _d], C.c_int),
    "chips_batch_destroy": ([_vp], C.c_int),
    "chips_demote_f64": ([_vp, _vp, _sz, _vp, _vp], C.c_int),
    "chips_fl_workspace_bytes": ([_i, _i], C.c_size_t),
    "chips_fl_candidates": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, C.POINTER(FlParams), _vp, _vp, _vp, _vp],
                            C.c_int),
    "chips_apply_candidates": ([_PP, _vp, _vp, _vp], C.c_int),
    "chips_pack_cube": ([_PP, _i, _vp, _vp, _vp], C.c_int),
    "chips_row_cosine": ([_vp, _vp, _i, _i, _i, _vp, _vp], C.c_int),
    "chips_input_rect": ([_PP, _vp], C.c_int),
    "chips_input_spans": ([_PP, _i, _i, _i, _vp, _i], C.c_int),
    "chips_copy_rect": ([_vp, _vp, _sz, _sz, _i, _sz, _i, _i, _vp], C.c_int),
    "chips_contours_workspace_bytes": ([_i, _i, C.c_longlong], C.c_size_t),
    "chips_contours": ([_vp, _i, _i, _i, _i, _i, C.c_longlong, _i, _vp, C.c_longlong, _vp, C.c_longlong,
                        _vp, _vp, _sz, _vp], C.c_int),
}

_lib = None


class ChipsNativeError(RuntimeError):
    pass


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise ChipsNativeError(
                f"{LIB} not found: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -m pychips_b200.build`.")
        lib = C.CDLL(LIB)
        for name, (args, ret) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if a declared symbol is missing
            fn.argtypes = args
            fn.restype = ret
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise ChipsNativeError(f"{what} failed with code {rc} "
                               "(-1 bad argument, -2 CUDA error, -3 workspace)")
