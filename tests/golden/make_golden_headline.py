"""Generate the HEADLINE-configuration fixtures from the UNMODIFIED reference (run in the BUILD
container, where `/root/reference` exists; no GPU needed):

  * `disk_n1024_k51`  — BASELINE config 1 (`/root/reference/test/test_method_flow.py:221`:
    medfilt_kernel=51, h_bins=500, threshold_range=[0,20]) at 1024^2: full arrays
    (filt_disk as float32, packed raw / cleaned maps, histogram, limits, probs).
  * `disk_n4096_k51`  — BASELINE config 2, 4096^2 with the reference defaults: sha256 digests of
    `filt_disk`, per-threshold digests of `raw_map` / `map`, pixel counts, the whole histogram,
    limits, probabilities, contour counts, plus a 64x64 grid of tile sums of `filt_disk` (so a
    digest mismatch can be localised without the 64 MB array).

The input image is NOT stored: `pychips_b200.synth.synthetic_disk(n, seed)` regenerates it and
the fixture carries its sha256.  For both cases the seconds each reference stage took on ONE
core of this container are recorded (`stage_seconds`) — the measured full-image CPU cost that
`bench.py`'s bounded `cpu_baseline` sample is validated against
(`profiles/r02_cpu_full_vs_sample.json`).  The oracle restatement (`oracle/chips_oracle.py`) is
run on the reference's own `filt_disk` for the stages after the median and asserted equal (the
median stage is the same `scipy.signal.medfilt2d` call; it is asserted equal at 1024^2).

    python tests/golden/make_golden_headline.py [1024|4096]
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import chips_oracle as co            # noqa: E402
from oracle import ref_harness as rh             # noqa: E402
from pychips_b200 import synth                   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KW = dict(medfilt_kernel=51, h_bins=500, threshold_range=[0, 20])


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def run_reference_timed(img32, r):
    """Stage methods of the unmodified `Chips` in `process_CHIPS` order, each one timed."""
    Chips, _ = rh.load_reference()
    d = synth.FakeDisk(img32.astype(np.float64), r)
    c = Chips(synth.FakeAIA(d), base_folder="/tmp/chips_ref/", **KW)
    secs = {}
    for name, fn in (("extract_solar_masks", c.extract_solar_masks), ("run_filters", c.run_filters),
                     ("extract_histogram", c.extract_histogram), ("extract_CHs_CHBs", c.extract_CHs_CHBs)):
        t = time.perf_counter()
        fn(d)
        secs[name] = time.perf_counter() - t
        print(f"  {name}: {secs[name]:.2f} s", flush=True)
    return d, secs


def check_oracle_after_median(d_ref, n, r):
    """The oracle restatement on the reference's own filt_disk (stages E-I)."""
    n_mask, i_mask = co.solar_masks(d_ref.raw.data, n, r)
    assert _same(n_mask, d_ref.solar_mask.n_mask) and _same(i_mask, d_ref.solar_mask.i_mask)
    filt = d_ref.solar_filter.filt_disk
    hd = co.extract_histogram(filt * n_mask, KW["h_bins"], None, 5)
    for f in ("hbase", "hbins", "bound", "h_thresh"):
        assert _same(getattr(d_ref.histogram, f), hd[f]), f
    assert _same(d_ref.histogram.peaks, hd["peaks"])
    regs = co.extract_CHs_CHBs(filt, n_mask, n, r, hd["peaks"], KW["threshold_range"])
    keys = list(d_ref.solar_ch_regions.__dict__.keys())
    assert keys == list(regs.keys())
    for k in keys:
        a, b = d_ref.solar_ch_regions.__dict__[k], regs[k]
        assert _same(a.map, b.map) and _same(a.prob, b.prob) and a.lim == b.lim, k
        assert len(a.contours) == len(b.contours)
    return regs


def common_fields(img, r, d, regs, secs):
    keys = list(d.solar_ch_regions.__dict__.keys())
    recs = [d.solar_ch_regions.__dict__[k] for k in keys]
    probs = [x.prob for x in recs]
    return keys, recs, dict(
        seed=0, pixel_radius=r, image_sha256=sha(img),
        hbase=d.histogram.hbase, hbins=d.histogram.hbins, bound=d.histogram.bound,
        h_thresh=d.histogram.h_thresh, peaks=np.array(d.histogram.peaks, dtype=np.float64),
        keys=np.array(keys), limits=np.array([x.lim for x in recs]), probs=np.array(probs),
        limit_0=recs[0].limit_0, n_contours=np.array([len(x.contours) for x in recs]),
        map_counts=np.array([int(x.map.sum()) for x in recs]),
        raw_counts=np.array([int(regs[k].raw_map.sum()) for k in keys]),
        kw_medfilt_kernel=KW["medfilt_kernel"], kw_h_bins=KW["h_bins"],
        kw_threshold_range=np.array(KW["threshold_range"]),
        stage_seconds=json.dumps(secs), cpu_info=json.dumps(dict(
            cores_used=1, note="one process, one thread, build container", nproc=os.cpu_count())))


def make_1024():
    n = 1024
    img, r = synth.synthetic_disk(n, 0)
    print("disk_n1024_k51: reference ...", flush=True)
    d, secs = run_reference_timed(img, r)
    d_or = synth.FakeDisk(img.astype(np.float64), r)
    co.OracleChips(**KW).run(d_or)
    assert _same(d.solar_filter.filt_disk, d_or.solar_filter.filt_disk)
    regs = check_oracle_after_median(d, n, r)
    keys, recs, out = common_fields(img, r, d, regs, secs)
    f32 = d.solar_filter.filt_disk.astype(np.float32)
    assert _same(f32.astype(np.float64), d.solar_filter.filt_disk)
    st, _ = co.prob_stack([x.map for x in recs], [x.prob for x in recs])
    out.update(filt_disk=f32, maps=np.packbits(np.array([x.map.astype(np.uint8) for x in recs])),
               raw_maps=np.packbits(np.array([regs[k].raw_map.astype(np.uint8) for k in keys])),
               stack_sha256=sha(st.astype(np.float64)))
    np.savez_compressed(os.path.join(HERE, "disk_n1024_k51.npz"), **out)
    print("disk_n1024_k51 written; total", sum(secs.values()), "s")


def make_4096():
    n = 4096
    img, r = synth.synthetic_disk(n, 0)
    print("disk_n4096_k51: reference (about 12 CPU-minutes) ...", flush=True)
    d, secs = run_reference_timed(img, r)
    regs = check_oracle_after_median(d, n, r)
    keys, recs, out = common_fields(img, r, d, regs, secs)
    f32 = d.solar_filter.filt_disk.astype(np.float32)
    assert _same(f32.astype(np.float64), d.solar_filter.filt_disk)
    tile = 64
    out.update(
        filt_sha256=sha(f32), i_mask_sha256=sha(np.packbits(d.solar_mask.i_mask.astype(np.uint8))),
        filt_tile_sums=f32.astype(np.float64).reshape(n // tile, tile, n // tile, tile).sum(axis=(1, 3)),
        map_sha256=np.array([sha(np.packbits(x.map.astype(np.uint8))) for x in recs]),
        raw_sha256=np.array([sha(np.packbits(regs[k].raw_map.astype(np.uint8))) for k in keys]))
    np.savez_compressed(os.path.join(HERE, "disk_n4096_k51.npz"), **out)
    # the full arrays stay outside the repo (debugging aid in the build container only)
    try:
        np.savez_compressed("/tmp/headline_4096_full.npz", filt_disk=f32,
                            maps=np.packbits(np.array([x.map.astype(np.uint8) for x in recs])),
                            raw_maps=np.packbits(np.array([regs[k].raw_map.astype(np.uint8) for k in keys])))
    except Exception as e:                                   # pragma: no cover
        print("could not keep the full arrays:", e)
    print("disk_n4096_k51 written; total", sum(secs.values()), "s")


if __name__ == "__main__":
    assert rh.available(), "reference tree missing"
    which = sys.argv[1:] or ["1024", "4096"]
    if "1024" in which:
        make_1024()
    if "4096" in which:
        make_4096()
