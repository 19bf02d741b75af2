"""Generate the committed golden fixtures (run in the BUILD container, where the reference
tree exists).  For every case: run the UNMODIFIED reference (oracle/ref_harness.py), run the
oracle restatement (oracle/chips_oracle.py) on the same input, assert they agree on every
stage output, and store the reference's outputs.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import chips_oracle as co            # noqa: E402
from oracle import ref_harness as rh             # noqa: E402
from pychips_b200 import synth                   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

DISK_CASES = {
    # name: (N, seed, scale, kwargs)
    "disk_n256_k11": (256, 0, 1.0, dict(medfilt_kernel=11, h_bins=500, threshold_range=[0, 20])),
    "disk_n256_k31_211": (256, 1, 1.0, dict(medfilt_kernel=31, h_bins=1000, threshold_range=[-3, 11])),
    "disk_n192_k51": (192, 2, 1.0, dict(medfilt_kernel=51, h_bins=500, threshold_range=[0, 20])),
    "disk_n256_k5_t28": (256, 3, 1.0, dict(medfilt_kernel=5, h_bins=5000, threshold_range=[-3, 25])),
    "disk_n200_k7_wiped": (200, 4, 0.02, dict(medfilt_kernel=7, h_bins=500, threshold_range=[-3, 5])),
    # List[float] threshold_range (chips.py:56): np.arange(-2.5, 6.25) = -2.5, -1.5, ..., 5.5
    "disk_n192_k11_frange": (192, 6, 1.0, dict(medfilt_kernel=11, h_bins=500, threshold_range=[-2.5, 6.25])),
}
# float32 raw / normalized arrays: numpy keeps float32 bin edges (np.histogram's bin_type), bound,
# limit_0 and the sigmoid of calculate_prob - the dtype real synoptic FITS maps come in
F32_DISK_CASES = {
    "disk_n256_k11_f32": (256, 7, 1.0, dict(medfilt_kernel=11, h_bins=500, threshold_range=[0, 20])),
    "disk_n192_k31_f32": (192, 8, 1.0, dict(medfilt_kernel=31, h_bins=5000, threshold_range=[-3, 11])),
}
F32_SYN_CASES = {
    "syn_180x360_k3_f32": (180, 360, 3, dict(medfilt_kernel=3, h_bins=5000, threshold_range=[-5, 20])),
}
SYN_CASES = {
    "syn_180x360_k3": (180, 360, 0, dict(medfilt_kernel=3, h_bins=5000, threshold_range=[-5, 20])),
    "syn_96x240_k11": (96, 240, 1, dict(medfilt_kernel=11, h_bins=500, threshold_range=[-5, 20])),
    "syn_90x200_k5_frange": (90, 200, 2, dict(medfilt_kernel=5, h_bins=1000, threshold_range=[0.1, 7.3])),
}


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def make_disk(name, n, seed, scale, kw, dtype=np.float64):
    img, r = synth.synthetic_disk(n, seed)
    img = (img * np.float32(scale)).astype(np.float32)
    d_ref = synth.FakeDisk(img.astype(dtype), r)
    rh.run_disk(d_ref, synth.FakeAIA(d_ref), **kw)
    d_or = synth.FakeDisk(img.astype(dtype), r)
    co.OracleChips(**kw).run(d_or)
    assert d_ref.histogram.hbins.dtype == dtype and d_ref.solar_filter.filt_disk.dtype == dtype
    assert _same(d_ref.solar_mask.n_mask, d_or.solar_mask.n_mask)
    assert _same(d_ref.solar_filter.filt_disk, d_or.solar_filter.filt_disk)
    assert _same(d_ref.solar_filter.solar_disk, d_or.solar_filter.solar_disk)
    for f in ("hbase", "hbins", "bound", "h_thresh"):
        assert _same(getattr(d_ref.histogram, f), getattr(d_or.histogram, f)), f
    assert _same(d_ref.histogram.peaks, d_or.histogram.peaks)
    out = dict(image=img, pixel_radius=r, i_mask=np.packbits(d_ref.solar_mask.i_mask.astype(np.uint8)),
               filt_disk=d_ref.solar_filter.filt_disk.astype(np.float32),
               hbase=d_ref.histogram.hbase, hbins=d_ref.histogram.hbins,
               bound=d_ref.histogram.bound, h_thresh=d_ref.histogram.h_thresh,
               peaks=np.array(d_ref.histogram.peaks, dtype=dtype),
               kw_medfilt_kernel=kw["medfilt_kernel"], kw_h_bins=kw["h_bins"],
               kw_threshold_range=np.array(kw["threshold_range"]), input_dtype=np.dtype(dtype).name)
    assert _same(out["filt_disk"].astype(np.float64), d_ref.solar_filter.filt_disk)
    has = hasattr(d_ref, "solar_ch_regions")
    assert has == hasattr(d_or, "solar_ch_regions")
    if has:
        keys = list(d_ref.solar_ch_regions.__dict__.keys())
        assert keys == list(d_or.solar_ch_regions.__dict__.keys())
        maps, probs, lims, ncont = [], [], [], []
        for k in keys:
            a, b = d_ref.solar_ch_regions.__dict__[k], d_or.solar_ch_regions.__dict__[k]
            assert _same(a.map, b.map) and _same(a.prob, b.prob) and a.lim == b.lim, k
            assert len(a.contours) == len(b.contours)
            maps.append(a.map.astype(np.uint8))
            probs.append(a.prob)
            lims.append(a.lim)
            ncont.append(len(a.contours))
        out.update(keys=np.array(keys), limits=np.array(lims), probs=np.array(probs),
                   n_contours=np.array(ncont), limit_0=d_ref.solar_ch_regions.__dict__[keys[0]].limit_0,
                   maps=np.packbits(np.array(maps)), raw_maps=np.packbits(
                       np.array([d_or.solar_ch_regions.__dict__[k].raw_map.astype(np.uint8) for k in keys])))
        try:       # plots.py:314 raises on an all-NaN probability list (nothing to stack)
            st, npb = co.prob_stack([d_ref.solar_ch_regions.__dict__[k].map for k in keys], probs)
            out.update(stack=st.astype(np.float64))
        except ValueError:
            pass
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "peaks", len(out["peaks"]), "regions", has, "T", len(out.get("limits", [])))


def make_syn(name, h, w, seed, kw, dtype=np.float64):
    img = synth.synthetic_synoptic(h, w, seed)
    s_ref = synth.FakeSynoptic(img.astype(dtype))
    rh.run_synoptic(s_ref, **kw)
    s_or = synth.FakeSynoptic(img.astype(dtype))
    co.OracleSynopticChips(**kw).run(s_or)
    assert s_ref.histogram.bound.dtype == dtype
    assert _same(s_ref.solar_filter, s_or.solar_filter)
    for f in ("hbase", "bound", "h_thresh"):
        assert _same(getattr(s_ref.histogram, f), getattr(s_or.histogram, f)), f
    assert _same(s_ref.histogram.peaks, s_or.histogram.peaks)
    keys = list(s_ref.solar_ch_regions.__dict__.keys())
    maps, probs, lims = [], [], []
    for k in keys:
        a, b = s_ref.solar_ch_regions.__dict__[k], s_or.solar_ch_regions.__dict__[k]
        assert _same(a.map, b.map) and _same(a.prob, b.prob), k
        maps.append(a.map.astype(np.uint8)); probs.append(a.prob); lims.append(a.lim)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), image=img, solar_filter=s_ref.solar_filter.astype(np.float32),
        hbase=s_ref.histogram.hbase, bound=s_ref.histogram.bound, h_thresh=s_ref.histogram.h_thresh,
        peaks=np.array(s_ref.histogram.peaks), keys=np.array(keys), limits=np.array(lims),
        probs=np.array(probs), limit_0=s_ref.solar_ch_regions.__dict__[keys[0]].limit_0,
        maps=np.packbits(np.array(maps)), kw_medfilt_kernel=kw["medfilt_kernel"],
        kw_h_bins=kw["h_bins"], kw_threshold_range=np.array(kw["threshold_range"]),
        input_dtype=np.dtype(dtype).name)
    print(name, "peaks", len(s_ref.histogram.peaks), "T", len(keys))


FL_CASES = {
    # name: (N, seed, kwargs of the 193 A detection that consumes the candidate bitmap)
    "fl_n256_k11": (256, 5, dict(medfilt_kernel=11, h_bins=500, threshold_range=[0, 12])),
}


def make_fl(name, n, seed, kw):
    """UNMODIFIED `chips.cleanup.CleanFl` (with skimage's resize stubbed by the identity, see
    oracle/cleanfl_oracle.py) and the UNMODIFIED Chips stage methods with `run_fl_cleanup`."""
    import types
    from argparse import Namespace
    from oracle import cleanfl_oracle as fo
    rh.load_reference()
    sk = types.ModuleType("skimage.transform")
    sk.resize = fo.identity_resize
    sys.modules["skimage.transform"] = sk
    sys.modules.pop("chips.cleanup", None)
    from chips.cleanup import CleanFl
    aia = synth.make_triplet_aia(n, seed)
    aia.save_dir = "/tmp/"
    imgs = {w: aia.datasets[w][n].normalized.data for w in (171, 193, 211)}
    r = aia.datasets[193][n].pixel_radius
    cf = CleanFl(aia, resolution=n)
    ret = cf.create_coronal_hole_candidates()
    o = fo.create_coronal_hole_candidates(imgs[171], imgs[193], imgs[211], r, n)
    for a, b in zip(ret, (o.candidate, o.bmmix, o.bmhot, o.bmcool, o.disk_mask)):
        assert _same(a, b)
    for a, b in ((cf.candidate_bitmap, o.candidate_bitmap), (cf.bmmix, o.bmmix_m),
                 (cf.bmhot, o.bmhot_m), (cf.bmcool, o.bmcool_m)):
        assert _same(a, b)
    # detection on the 193 A disk with the hook chips.py:376-377
    Chips, _ = rh.load_reference()
    d_ref = aia.datasets[193][n]
    c = Chips(aia, base_folder="/tmp/chips_ref/", run_fl_cleanup=True, **kw)
    c.cf = cf
    c.extract_solar_masks(d_ref); c.run_filters(d_ref); c.extract_histogram(d_ref); c.extract_CHs_CHBs(d_ref)
    d_or = synth.FakeDisk(imgs[193].copy(), r)
    co.OracleChips(candidate_bitmap=o.candidate_bitmap, **kw).run(d_or)
    keys = list(d_ref.solar_ch_regions.__dict__.keys())
    assert keys == list(d_or.solar_ch_regions.__dict__.keys())
    maps, raw, probs, lims = [], [], [], []
    for k in keys:
        a, b = d_ref.solar_ch_regions.__dict__[k], d_or.solar_ch_regions.__dict__[k]
        assert _same(a.map, b.map) and _same(a.prob, b.prob), k
        maps.append(a.map.astype(np.uint8)); raw.append(b.raw_map.astype(np.uint8))
        probs.append(a.prob); lims.append(a.lim)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), seed=seed, n=n, pixel_radius=r,
        img171=imgs[171].astype(np.float32), img193=imgs[193].astype(np.float32),
        img211=imgs[211].astype(np.float32),
        planes=np.packbits(np.array([ret[0], ret[1], ret[2], ret[3], cf.candidate_bitmap]).astype(np.uint8)),
        thresholds=np.array(o.thresholds), means=np.array(o.means),
        keys=np.array(keys), limits=np.array(lims), probs=np.array(probs),
        maps=np.packbits(np.array(maps)), raw_maps=np.packbits(np.array(raw)),
        kw_medfilt_kernel=kw["medfilt_kernel"], kw_h_bins=kw["h_bins"],
        kw_threshold_range=np.array(kw["threshold_range"]))
    on = o.disk_mask > 0
    print(name, "candidate fill on disk", float(o.candidate[on].mean()), "T", len(keys))


def make_netcdf_layout(name="netcdf_layout_n256", src="disk_n256_k11"):
    """The NETCDF4 group layout of the UNMODIFIED reference's Chips.to_netcdf (chips.py:512-590),
    recorded call by call (netCDF4 itself is absent from this image: oracle/nc_recorder.py)."""
    import json
    import sys as _sys
    from oracle import nc_recorder as ncr
    n, seed, scale, kw = DISK_CASES[src]
    img, r = synth.synthetic_disk(n, seed)
    img = (img * np.float32(scale)).astype(np.float32)
    d = synth.FakeDisk(img.astype(np.float64), r)
    aia = synth.FakeAIA(d)
    c = rh.run_disk(d, aia, **kw)
    mod = _sys.modules[type(c).__module__]
    mod.Dataset = ncr.Dataset                      # the name chips.py:526 `nc = Dataset(...)` resolves
    c.to_netcdf(193, n)
    log = ncr.LAST[-1].log
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(dict(source=src, n=n, log=log), f, indent=1)
    print(name, len(log), "calls:", [x[:3] for x in log if x[0] in ("group", "var")])


if __name__ == "__main__":
    assert rh.available(), "reference tree missing"
    if len(sys.argv) > 1 and sys.argv[1] == "netcdf":
        make_netcdf_layout()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fl":
        for name, (n, seed, kw) in FL_CASES.items():
            make_fl(name, n, seed, kw)
        sys.exit(0)
    only = set(sys.argv[1:])             # optional: names of the cases to (re)generate
    for name, (n, seed, scale, kw) in DISK_CASES.items():
        if not only or name in only:
            make_disk(name, n, seed, scale, kw)
    for name, (h, w, seed, kw) in SYN_CASES.items():
        if not only or name in only:
            make_syn(name, h, w, seed, kw)
    for name, (n, seed, kw) in FL_CASES.items():
        if not only or name in only:
            make_fl(name, n, seed, kw)
    for name, (n, seed, scale, kw) in F32_DISK_CASES.items():
        if not only or name in only:
            make_disk(name, n, seed, scale, kw, np.float32)
    for name, (h, w, seed, kw) in F32_SYN_CASES.items():
        if not only or name in only:
            make_syn(name, h, w, seed, kw, np.float32)
