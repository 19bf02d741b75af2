"""TEST/BENCH INFRASTRUCTURE — times the CPU oracle (reference control flow + the same
scipy / numpy / OpenCV calls, `oracle/chips_oracle.py`) on a BOUNDED sample of one full-disk
detection and extrapolates to the whole image.

One 4096^2 detection at the reference default k=51 costs ~11 CPU-minutes
(BASELINE.md §2), so `bench.py` times:
  * the median filter (chips.py:214-216) on a horizontal band of `band_rows` output rows through
    the disk centre (full width).  scipy.signal.medfilt2d does the same k^2 quick-select for
    every output pixel, padded border included, so the cost is linear in output pixels and the
    band time is scaled by N^2 / band pixels;
  * masks + histogram at full size, and the `for lim in limits` loop (chips.py:360-400) for the
    first `thr_sample` of the T thresholds, scaled by T / thr_sample (iterations are
    independent and do identical work).  The loop needs a filtered image: the caller passes the
    real `filt_disk` of the same image when it has one (bench.py's `cpu_baseline` leg copies the
    GPU result back BEFORE the timed CPU region; it is only an input of the timed numpy / OpenCV
    calls).  The `--impl reference` arm may not touch the GPU path at all and uses a stand-in
    with the same structure (median of the 4x decimated image, pixel-replicated back).

How good the extrapolation is was checked against the full-image run of the UNMODIFIED reference
(tests/golden/make_golden_headline.py records its stage times in the 4096^2 fixture):
`profiles/r02_cpu_full_vs_sample.json`, written by `python -m oracle.cpu_timing validate`.
"""
from __future__ import annotations

import os
import time

import numpy as np
from scipy import signal

from . import chips_oracle as co


def proxy_filtered(img64: np.ndarray, k: int) -> np.ndarray:
    f = 4
    small = img64[::f, ::f]
    ks = max(3, (k // f) | 1)
    med = signal.medfilt2d(np.ascontiguousarray(small), ks)
    return np.repeat(np.repeat(med, f, axis=0), f, axis=1)[:img64.shape[0], :img64.shape[1]]


def sample_detection_seconds(img32: np.ndarray, pixel_radius: int, k: int = 51, h_bins: int = 500,
                             threshold_range=(0, 20), band_rows: int = 16, thr_sample: int = 2,
                             filt: np.ndarray | None = None):
    """Returns dict(seconds_per_image=extrapolated, sample_seconds=measured, parts=...)."""
    n = img32.shape[0]
    img = img32.astype(np.float64)
    half = k // 2
    y0 = n // 2 - band_rows // 2
    lo, hi = max(0, y0 - half), min(n, y0 + band_rows + half)
    band = np.ascontiguousarray(img[lo:hi])
    t = time.perf_counter()
    signal.medfilt2d(band, k)
    t_band = time.perf_counter() - t
    t_median = t_band * (n * n) / float(band.shape[0] * band.shape[1])

    t = time.perf_counter()
    n_mask, i_mask = co.solar_masks(img, n, pixel_radius)
    t_masks = time.perf_counter() - t
    real = filt is not None
    filt = (np.asarray(filt, dtype=np.float64) if real else proxy_filtered(img, k)) * i_mask
    t = time.perf_counter()
    hd = co.extract_histogram(filt * n_mask, h_bins, None, 5)
    t_hist = time.perf_counter() - t
    T = int(threshold_range[1] - threshold_range[0])
    m = max(1, min(thr_sample, T))
    t = time.perf_counter()
    co.extract_CHs_CHBs(filt, n_mask, n, pixel_radius, hd["peaks"],
                        [threshold_range[0], threshold_range[0] + m], keep_contours=False)
    t_loop = time.perf_counter() - t
    t_rest = t_masks + t_hist + t_loop * (T / m)
    return dict(seconds_per_image=t_median + t_rest,
                sample_seconds=t_band + t_masks + t_hist + t_loop,
                parts=dict(median=t_median, masks=t_masks, histogram=t_hist, regions=t_loop * T / m),
                sample=f"medfilt2d on a {band.shape[0]}x{band.shape[1]} band ({band_rows} output rows, "
                       f"scaled x{n * n / float(band.shape[0] * band.shape[1]):.1f}) + masks + "
                       f"histogram at full size + {m} of {T} thresholds (scaled x{T / m:.1f}) on "
                       + ("the real filt_disk of this image" if real else "a decimated-median stand-in"))


def _worker(args):
    n, seed, k, h_bins, tr, band_rows, thr_sample = args
    from pychips_b200 import synth
    img, r = synth.synthetic_disk(n, seed)
    return sample_detection_seconds(img, r, k, h_bins, tr, band_rows, thr_sample)


def pool_throughput(n: int, k: int, h_bins: int, tr, processes: int, band_rows: int = 16,
                    thr_sample: int = 2, pool=None):
    """`processes` independent images in parallel, one per process (the reference itself is
    single-threaded).  Returns (images_per_second, per-image dicts)."""
    import multiprocessing as mp
    args = [(n, s, k, h_bins, tuple(tr), band_rows, thr_sample) for s in range(processes)]
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(processes)
    try:
        outs = pool.map(_worker, args)
    finally:
        if own:
            pool.close()
            pool.join()
    per_image = float(np.mean([o["seconds_per_image"] for o in outs]))
    return processes / per_image, outs


def usable_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def validate(out_path: str | None = None, band_rows: int = 16, thr_sample: int = 2):
    """Build container only: the bounded sample's extrapolation against the measured full-image
    run of the unmodified reference on the SAME machine (stage times stored in the 4096^2 fixture
    by tests/golden/make_golden_headline.py).  Writes profiles/r02_cpu_full_vs_sample.json."""
    import json
    from pychips_b200 import synth
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "disk_n4096_k51.npz"))
    secs = json.loads(str(g["stage_seconds"]))
    img, r = synth.synthetic_disk(4096, int(g["seed"]))
    full = "/tmp/headline_4096_full.npz"
    filt = np.load(full)["filt_disk"] if os.path.exists(full) else None
    rows = []
    for f, label in ((filt, "real filt_disk"), (None, "stand-in")):
        if f is None and label == "real filt_disk":
            continue
        o = sample_detection_seconds(img, r, 51, 500, (0, 20), band_rows, thr_sample, filt=f)
        rows.append(dict(threshold_loop_input=label, extrapolated_s=round(o["seconds_per_image"], 1),
                         sample_s=round(o["sample_seconds"], 2), parts={k: round(v, 2) for k, v in o["parts"].items()}))
    measured = float(sum(secs.values()))
    doc = dict(workload="4096^2, k=51, h_bins=500, threshold_range=[0,20], synthetic_disk(4096, 0)",
               machine="build container, one core (taskset), idle otherwise",
               measured_full_image_s=round(measured, 1), measured_stage_seconds=secs,
               samples=rows,
               extrapolated_over_measured=[round(x["extrapolated_s"] / measured, 3) for x in rows],
               measured_parts=dict(median=secs["run_filters"], regions=secs["extract_CHs_CHBs"]))
    out_path = out_path or os.path.join(root, "profiles", "r02_cpu_full_vs_sample.json")
    with open(out_path, "w") as fh:
        json.dump(doc, fh, indent=1)
    return doc


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "validate":
        print(validate())
