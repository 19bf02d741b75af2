"""CPU: the oracle restatement (oracle/chips_oracle.py) against the committed outputs of the
UNMODIFIED reference (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import chips_oracle as co
from pychips_b200 import synth

DISK = ["disk_n256_k11", "disk_n256_k31_211", "disk_n192_k51", "disk_n256_k5_t28",
        "disk_n200_k7_wiped", "disk_n192_k11_frange", "disk_n256_k11_f32", "disk_n192_k31_f32"]
SYN = ["syn_180x360_k3", "syn_96x240_k11", "syn_90x200_k5_frange", "syn_180x360_k3_f32"]


def in_dtype(g):
    """dtype of the arrays the reference was given (float32 goldens: numpy's float32 histogram)."""
    return np.dtype(str(g["input_dtype"])) if "input_dtype" in g else np.dtype(np.float64)


def unpack(bits, T, H, W):
    return np.unpackbits(bits)[:T * H * W].reshape(T, H, W)


@pytest.mark.parametrize("name", DISK)
def test_disk_oracle_matches_reference_golden(name):
    g = load_golden(name)
    img = g["image"]
    n = img.shape[0]
    d = synth.FakeDisk(img.astype(in_dtype(g)), int(g["pixel_radius"]))
    o = co.OracleChips(medfilt_kernel=int(g["kw_medfilt_kernel"]), h_bins=int(g["kw_h_bins"]),
                       threshold_range=list(g["kw_threshold_range"]))
    o.run(d)
    assert d.histogram.hbins.dtype == g["hbins"].dtype == in_dtype(g)
    assert np.array_equal(np.unpackbits(g["i_mask"])[:n * n].reshape(n, n),
                          d.solar_mask.i_mask.astype(np.uint8))
    assert np.array_equal(g["filt_disk"].astype(np.float64), d.solar_filter.filt_disk)
    assert np.array_equal(g["hbase"], d.histogram.hbase)
    assert np.array_equal(g["hbins"], d.histogram.hbins)
    assert np.array_equal(g["bound"], d.histogram.bound)
    assert g["h_thresh"] == d.histogram.h_thresh
    assert np.array_equal(g["peaks"], np.array(d.histogram.peaks))
    keys = list(d.solar_ch_regions.__dict__.keys())
    assert keys == list(g["keys"])
    T = len(keys)
    maps = unpack(g["maps"], T, n, n)
    raw = unpack(g["raw_maps"], T, n, n)
    for t, k in enumerate(keys):
        r = d.solar_ch_regions.__dict__[k]
        assert np.array_equal(maps[t], r.map.astype(np.uint8))
        assert np.array_equal(raw[t], r.raw_map.astype(np.uint8))
        assert np.array_equal(g["probs"][t], r.prob, equal_nan=True)
        assert g["limits"][t] == r.lim
        assert g["n_contours"][t] == len(r.contours)


@pytest.mark.parametrize("name", SYN)
def test_synoptic_oracle_matches_reference_golden(name):
    g = load_golden(name)
    img = g["image"]
    s = synth.FakeSynoptic(img.astype(in_dtype(g)))
    o = co.OracleSynopticChips(medfilt_kernel=int(g["kw_medfilt_kernel"]),
                               h_bins=int(g["kw_h_bins"]),
                               threshold_range=list(g["kw_threshold_range"]))
    o.run(s)
    assert np.array_equal(g["solar_filter"].astype(np.float64), s.solar_filter)
    assert np.array_equal(g["hbase"], s.histogram.hbase)
    assert np.array_equal(g["peaks"], np.array(s.histogram.peaks))
    keys = list(s.solar_ch_regions.__dict__.keys())
    assert keys == list(g["keys"])
    T, (H, W) = len(keys), img.shape
    maps = unpack(g["maps"], T, H, W)
    for t, k in enumerate(keys):
        r = s.solar_ch_regions.__dict__[k]
        assert np.array_equal(maps[t], r.map.astype(np.uint8))
        assert np.array_equal(g["probs"][t], r.prob, equal_nan=True)


def test_reference_still_agrees_when_present():
    """Where the reference tree exists (build container) re-run it live on one case."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present (GPU box)")
    img, r = synth.synthetic_disk(160, 11)
    d1 = synth.FakeDisk(img.astype(np.float64), r)
    rh.run_disk(d1, synth.FakeAIA(d1), medfilt_kernel=9, h_bins=300, threshold_range=[-2, 9])
    d2 = synth.FakeDisk(img.astype(np.float64), r)
    co.OracleChips(medfilt_kernel=9, h_bins=300, threshold_range=[-2, 9]).run(d2)
    assert np.array_equal(d1.solar_filter.filt_disk, d2.solar_filter.filt_disk)
    assert np.array_equal(d1.histogram.hbase, d2.histogram.hbase)
    assert list(d1.solar_ch_regions.__dict__) == list(d2.solar_ch_regions.__dict__)
    for k in d1.solar_ch_regions.__dict__:
        a, b = d1.solar_ch_regions.__dict__[k], d2.solar_ch_regions.__dict__[k]
        assert np.array_equal(a.map, b.map)
        assert np.array_equal(a.prob, b.prob, equal_nan=True)


def test_cleanfl_oracle_matches_reference_golden():
    """oracle/cleanfl_oracle.py + the candidate hook of chips_oracle against the outputs of the
    unmodified `chips.cleanup.CleanFl` / `Chips(run_fl_cleanup=True)` stage methods."""
    from oracle import cleanfl_oracle as fo
    g = load_golden("fl_n256_k11")
    n, r = int(g["n"]), int(g["pixel_radius"])
    imgs = [g[k].astype(np.float64) for k in ("img171", "img193", "img211")]
    o = fo.create_coronal_hole_candidates(imgs[0], imgs[1], imgs[2], r, n)
    planes = np.unpackbits(g["planes"])[:5 * n * n].reshape(5, n, n)
    for i, a in enumerate((o.candidate, o.bmmix, o.bmhot, o.bmcool, o.candidate_bitmap)):
        assert np.array_equal(planes[i], a.astype(np.uint8)), i
    assert np.array_equal(g["thresholds"], np.array(o.thresholds))
    assert np.array_equal(g["means"], np.array(o.means))
    d = synth.FakeDisk(imgs[1], r)
    co.OracleChips(medfilt_kernel=int(g["kw_medfilt_kernel"]), h_bins=int(g["kw_h_bins"]),
                   threshold_range=list(g["kw_threshold_range"]),
                   candidate_bitmap=o.candidate_bitmap).run(d)
    keys = list(d.solar_ch_regions.__dict__.keys())
    assert keys == list(g["keys"])
    T = len(keys)
    maps, raw = unpack(g["maps"], T, n, n), unpack(g["raw_maps"], T, n, n)
    for t, k in enumerate(keys):
        reg = d.solar_ch_regions.__dict__[k]
        assert np.array_equal(maps[t], reg.map.astype(np.uint8))
        assert np.array_equal(raw[t], reg.raw_map.astype(np.uint8))
        assert np.array_equal(g["probs"][t], reg.prob, equal_nan=True)


def test_headline_config1_oracle_matches_reference_golden():
    """BASELINE config 1 (1024^2, k=51, h_bins=500, [0,20]): the oracle's stages after the median on
    the reference's own filt_disk, and the median itself on 48 sampled windows (the full
    scipy.signal.medfilt2d call costs 29 s; make_golden_headline.py asserted it in full)."""
    import hashlib
    g = load_golden("disk_n1024_k51")
    n, r = 1024, int(g["pixel_radius"])
    img, r2 = synth.synthetic_disk(n, int(g["seed"]))
    assert r2 == r
    if hashlib.sha256(img.tobytes()).hexdigest() != str(g["image_sha256"]):
        pytest.skip("synthetic generator not bit-reproducible on this host")
    filt = g["filt_disk"].astype(np.float64)
    n_mask, i_mask = co.solar_masks(img.astype(np.float64), n, r)
    rng = np.random.default_rng(0)
    ys, xs = np.nonzero(i_mask)
    pad = np.pad(img, 25)
    for i in rng.integers(0, len(ys), 48):
        w = np.sort(pad[ys[i]:ys[i] + 51, xs[i]:xs[i] + 51].ravel())
        assert filt[ys[i], xs[i]] == w[1300]
    assert np.all(filt[i_mask == 0] == 0)
    hd = co.extract_histogram(filt * n_mask, 500, None, 5)
    for f in ("hbase", "hbins", "bound", "h_thresh"):
        assert np.array_equal(g[f], hd[f]), f
    assert np.array_equal(g["peaks"], np.array(hd["peaks"]))
    regs = co.extract_CHs_CHBs(filt, n_mask, n, r, hd["peaks"], [0, 20])
    assert list(regs.keys()) == list(g["keys"])
    T = len(regs)
    maps, raw = unpack(g["maps"], T, n, n), unpack(g["raw_maps"], T, n, n)
    for t, (k, reg) in enumerate(regs.items()):
        assert np.array_equal(maps[t], reg.map.astype(np.uint8))
        assert np.array_equal(raw[t], reg.raw_map.astype(np.uint8))
        assert g["probs"][t] == reg.prob and g["limits"][t] == reg.lim
        assert g["n_contours"][t] == len(reg.contours)
        assert g["map_counts"][t] == int(reg.map.sum()) and g["raw_counts"][t] == int(reg.raw_map.sum())


def test_headline_config2_fixture_is_self_consistent():
    """The 4096^2 digest fixture: keys / limits / histogram are consistent with each other and the
    recorded reference stage times are there (bench.py validates its bounded CPU sample with them)."""
    import json
    import os
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, "disk_n4096_k51.npz")):
        pytest.skip("fixture not generated yet")
    g = load_golden("disk_n4096_k51")
    assert len(g["keys"]) == 20 and len(g["map_sha256"]) == 20 and len(g["raw_sha256"]) == 20
    assert np.array_equal(g["limits"], np.round(g["limit_0"] + np.arange(0, 20), 4))
    assert [str(v) for v in g["limits"]] == list(g["keys"])
    assert g["limit_0"] == g["peaks"][0]
    assert np.array_equal(g["bound"], g["hbins"][:-1] + np.diff(g["hbins"]))
    assert np.all(np.diff(g["raw_counts"]) >= 0) and np.all(g["map_counts"] >= 0)
    secs = json.loads(str(g["stage_seconds"]))
    assert set(secs) == {"extract_solar_masks", "run_filters", "extract_histogram", "extract_CHs_CHBs"}
    assert secs["run_filters"] > 60


def test_sliding_windows_restatement_is_the_regional_one_without_overlap():
    """oracle.sliding_histograms (parity unpinned: the reference method is a stub) with overlap 0 on
    square splits visits the regions of oracle.extract_histograms in the same order with the same
    outputs; with overlap the windows tile the image with the stated step."""
    from pychips_b200 import synth
    img, r = synth.synthetic_disk(256, 3)
    d = synth.FakeDisk(img.astype(np.float64), r)
    o = co.OracleChips(medfilt_kernel=5, h_bins=200)
    o.run(d, keep_contours=False)
    a, fa, ha = co.extract_histograms(d.solar_filter.filt_disk, d.solar_mask.n_mask, 256, 2, 2, 200, None, 5)
    b, fb, hb = co.sliding_histograms(d.solar_filter.filt_disk, d.solar_mask.n_mask, 256, 2, 2, 200, None, 5, 0.0)
    assert len(a) == len(b) == 4 and ha == hb and [int(v) for v in fa] == [int(v) for v in fb]
    for k in a:
        assert np.array_equal(a[k].hbase, b[k].hbase, equal_nan=True) and np.array_equal(a[k].bound, b[k].bound)
        assert np.array_equal(a[k].peaks, b[k].peaks)
    w = co.sliding_windows(256, 2, 2, 0.5)
    assert len(w) == 9 and w[0] == (0, 128, 0, 128) and w[1] == (0, 128, 64, 192) and w[-1] == (128, 256, 128, 256)
