"""GPU parity tests: the CUDA path (through the C ABI / the drop-in classes) against the
oracle, the committed reference goldens and, at full size, an independent brute-force GPU
selection plus size-independent properties.  Bit-exact for medians, counts, masks, labels;
1e-5 relative for thresholds / probabilities (they are in fact compared exactly too)."""
import numpy as np
import pytest
import torch
from scipy import ndimage, signal

from conftest import load_golden
from oracle import chips_oracle as co
from oracle import first_principles as fp
from pychips_b200 import synth

pytestmark = pytest.mark.gpu

REL = 1e-5


def _engine():
    from pychips_b200 import engine
    return engine


def gpu_medfilt(img, k, cx=0, cy=0, r=-1, brute=False):
    eng = _engine()
    from pychips_b200 import _lib
    dev = eng.require_cuda()
    t = torch.from_numpy(np.ascontiguousarray(img)).to(dev)
    out = torch.full_like(t, -777.0)
    H, W = img.shape
    if brute:
        lib = _lib.load()
        _lib.check(lib.chips_medfilt2d_bruteforce(eng._ptr(t), eng._ptr(out), H, W, k,
                                                  t.element_size(), cx, cy, r, eng._stream_ptr()),
                   "brute")
        return out.cpu().numpy(), None
    det = eng.Detector(eng.Geometry(H, W, cx, cy, r), k, 8, 0, 1, t.element_size(), dev)
    det.medfilt(t, out)
    res = det.result()
    return out.cpu().numpy(), res


# ----------------------------------------------------------------------------- median
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("k", [3, 5, 11, 31, 51])
def test_median_bruteforce_kernel_vs_scipy(k, dtype):
    rng = np.random.default_rng(k)
    img = rng.gamma(2.0, 30.0, size=(70, 93)).astype(dtype)
    img[10:14, 20:50] = 0.0
    img[30:33, :] = -2.5
    got, _ = gpu_medfilt(img, k, brute=True)
    assert np.array_equal(got, signal.medfilt2d(img, k))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("k", [1, 3, 5, 7, 11, 31, 51, 71])
def test_median_fast_vs_scipy_unmasked(k, dtype):
    rng = np.random.default_rng(100 + k)
    img = rng.gamma(2.0, 30.0, size=(150, 203)).astype(dtype)
    img[10:40, 20:90] = 0.0                 # ties with the zero padding
    img[60:64, :] = -2.5                    # negatives
    img[100:140, 100:160] = 7.25            # constant block
    got, res = gpu_medfilt(img, k)
    assert res.median_errors == 0
    ref = signal.medfilt2d(img, k)
    bad = np.argwhere(got != ref)
    assert bad.size == 0, (len(bad), bad[:5], got[tuple(bad[0])], ref[tuple(bad[0])])


@pytest.mark.parametrize("k", [11, 31, 51])
def test_median_tma_tile_loads_and_plain_load_fallback_vs_scipy(k):
    # W % 4 == 0, a 16-byte aligned base and an image larger than a block's box: k_block_sort stages
    # every region with one TMA tile load whose out-of-image zero fill is medfilt2d's padding (all
    # four borders are met).  The same pixels at a base address that is NOT 16-byte aligned cannot be
    # described by a tensor map and take the plain-load path.
    eng = _engine()
    dev = eng.require_cuda()
    rng = np.random.default_rng(900 + k)
    H, W = 300, 512
    img = rng.gamma(2.0, 30.0, size=(H, W)).astype(np.float32)
    img[:40, :60] = 0.0                       # ties with the zero padding in a corner
    img[250:, 400:] = -3.0                    # negatives against the padding
    ref = signal.medfilt2d(img, k)
    got, res = gpu_medfilt(img, k)
    assert res.median_errors == 0 and np.array_equal(got, ref)
    flat = torch.empty(H * W + 1, dtype=torch.float32, device=dev)
    t = flat[1:].view(H, W)
    t.copy_(torch.from_numpy(img))
    assert t.data_ptr() % 16 != 0
    out = torch.full((H, W), -777.0, dtype=torch.float32, device=dev)
    det = eng.Detector(eng.Geometry(H, W, 0, 0, -1), k, 8, 0, 1, 4, dev)
    det.medfilt(t, out)
    assert det.result().median_errors == 0 and np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("shape,k", [((20, 20), 31), ((9, 300), 11), ((300, 9), 11), ((1, 1), 3),
                                     ((64, 64), 51), ((17, 33), 5)])
def test_median_fast_ragged_shapes(shape, k):
    rng = np.random.default_rng(shape[0] * 7 + k)
    img = rng.normal(50, 20, size=shape).astype(np.float32)
    got, res = gpu_medfilt(img, k)
    assert res.median_errors == 0
    assert np.array_equal(got, signal.medfilt2d(img, k))


def test_median_fast_degenerate_values():
    for img in (np.zeros((90, 90), np.float32), np.full((90, 90), 3.5, np.float32),
                np.tile(np.arange(90, dtype=np.float32), (90, 1)),
                (np.arange(8100).reshape(90, 90) % 2).astype(np.float32)):
        got, res = gpu_medfilt(img, 11)
        assert res.median_errors == 0
        assert np.array_equal(got, signal.medfilt2d(img, 11))


@pytest.mark.parametrize("n,k", [(256, 11), (256, 51), (200, 31)])
def test_median_fast_masked_vs_oracle(n, k):
    img, r = synth.synthetic_disk(n, 5)
    c = int(n / 2)
    got, res = gpu_medfilt(img, k, c, c, r)
    assert res.median_errors == 0
    _, i_mask = co.solar_masks(img.astype(np.float64), n, r)
    _, ref = co.run_filters(img.astype(np.float64), i_mask, k)
    assert np.array_equal(got.astype(np.float64), ref)


@pytest.mark.parametrize("n,k", [(1024, 51), (4096, 11), (4096, 51)])
def test_median_fast_vs_bruteforce_full_size(n, k):
    img, r = synth.synthetic_disk(n, 1)
    c = int(n / 2)
    fast, res = gpu_medfilt(img, k, c, c, r)
    assert res.median_errors == 0
    slow, _ = gpu_medfilt(img, k, c, c, r, brute=True)
    assert np.array_equal(fast, slow), int((fast != slow).sum())
    # spot-check 64 windows against a plain numpy median of the zero-padded window
    rng = np.random.default_rng(0)
    h = k // 2
    pad = np.pad(img, h)
    on = fp.circle_mask(n, n, c, c, r)
    ys, xs = np.nonzero(on)
    for i in rng.integers(0, len(ys), 64):
        y, x = ys[i], xs[i]
        w = np.sort(pad[y:y + k, x:x + k].ravel())
        assert fast[y, x] == w[(k * k - 1) // 2]
    assert np.all(fast[~on] == 0)


@pytest.mark.parametrize("q,k,dtype", [(0.5, 31, np.float32), (4.0, 51, np.float32), (1.0, 11, np.float64),
                                       (25.0, 51, np.float64)])
def test_median_fast_quantised_intensities(q, k, dtype):
    """Instrument-like quantised data: heavy duplicates (ties are broken by the block sort, every
    pixel still has a unique rank).  Bit-exact against scipy and the brute-force kernel."""
    img, r = synth.synthetic_disk(600, 5)
    img = (np.round(img / q) * q).astype(dtype)
    got, res = gpu_medfilt(img, k)
    assert res.median_errors == 0
    assert np.array_equal(got, signal.medfilt2d(img, k))
    c = 300
    fast, res = gpu_medfilt(img, k, c, c, r)
    slow, _ = gpu_medfilt(img, k, c, c, r, brute=True)
    assert res.median_errors == 0 and np.array_equal(fast, slow)


@pytest.mark.parametrize("k", [11, 51])
def test_median_float64_keys_that_agree_in_their_top_32_bits(k):
    """float64 storage: the block sort first orders the top 32 significant key bits only and
    re-sorts a block on all its bits when that order is not exact.  Values that differ in the
    last mantissa bits only, next to zeros and a huge outlier (so that the 32-bit window sits far
    above those bits), force the second attempt; plain noise takes the first."""
    rng = np.random.default_rng(k)
    close = 1.0 + rng.integers(0, 1 << 20, size=(180, 210)).astype(np.float64) * 2.0 ** -52
    img = close.copy()
    img[rng.random(img.shape) < 0.2] = 0.0
    img[7, 9] = 1e9
    assert len(np.unique(img)) > 10000 and not np.array_equal(img.astype(np.float32), img)
    for a in (img, rng.normal(100.0, 30.0, size=(180, 210))):
        got, res = gpu_medfilt(a, k)
        assert res.median_errors == 0
        assert np.array_equal(got, signal.medfilt2d(a, k))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("k", [7, 31, 73])
def test_median_flat_two_valued_and_maximum_kernel(k, dtype):
    """Inputs that used to take the overflow paths of the tile refinement (every pixel of a region
    in one bucket): constant, two-valued, a step edge; and the largest supported kernel (73)."""
    rng = np.random.default_rng(5 * k)
    flat = np.full((140, 170), 12.5, dtype)
    two = np.where(rng.random((140, 170)) < 0.5, 3.0, 4.0).astype(dtype)
    step = np.zeros((140, 170), dtype)
    step[:, 85:] = 1.0 + (rng.random((140, 85)) * (1e-3 if dtype == np.float64 else 0)).astype(dtype)
    for img in (flat, two, step):
        got, res = gpu_medfilt(img, k)
        assert res.median_errors == 0
        assert np.array_equal(got, signal.medfilt2d(img, k))
        got, res = gpu_medfilt(img, k, 85, 70, 60)
        ref = signal.medfilt2d(img, k)
        yy, xx = np.mgrid[:140, :170]
        on = (xx - 85) ** 2 + (yy - 70) ** 2 <= 60 ** 2
        assert res.median_errors == 0 and np.array_equal(got[on], ref[on]) and not got[~on].any()


def test_median_rejects_kernels_beyond_the_block_sort():
    from pychips_b200 import _lib
    import ctypes as C
    p = _lib.Plan()
    lib = _lib.load()
    assert lib.chips_plan_init(C.byref(p), 512, 512, 75, 8, 1, 4, 0, 0, -1) == -1
    assert lib.chips_plan_init(C.byref(p), 512, 512, 73, 8, 1, 8, 0, 0, -1) == 0 and p.blk_rows == 16


def test_median_quantised_full_size_is_not_slower():
    """4096^2, k=51 on 1-DN quantised data: exact (brute-force kernel) and no slower than on
    continuous data (without point buckets this case was 5x slower)."""
    img, r = synth.synthetic_disk(4096, 2)
    qimg = np.round(img).astype(np.float32)
    c = 2048
    fast, res = gpu_medfilt(qimg, 51, c, c, r)
    slow, _ = gpu_medfilt(qimg, 51, c, c, r, brute=True)
    assert res.median_errors == 0 and np.array_equal(fast, slow)
    eng = _engine()
    dev = eng.require_cuda()
    det = eng.Detector(eng.Geometry(4096, 4096, c, c, r), 51, 8, 0, 1, 4, dev)

    def ms(a):
        t = torch.from_numpy(a).to(dev)
        det.medfilt(t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            det.medfilt(t)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3
    assert ms(qimg) < 1.5 * ms(img)


# ----------------------------------------------------------------------------- full pipeline
def in_dtype(g):
    """dtype of the arrays the reference was given (float32 goldens: numpy's float32 histogram)."""
    return np.dtype(str(g["input_dtype"])) if "input_dtype" in g else np.dtype(np.float64)


def run_gpu_disk(img, r, dtype=np.float64, **kw):
    from pychips_b200 import Chips
    d = synth.FakeDisk(img.astype(dtype), r)
    c = Chips(synth.FakeAIA(d), base_folder="/tmp/chips_b200_test/", **kw)
    c.run_CHIPS()
    return c, d


def compare_disk(d, g_filt, g_hbase, g_hbins, g_bound, g_peaks, g_lims, g_probs, g_raw, g_maps,
                 g_stack=None, chips=None):
    assert np.array_equal(d.solar_filter.filt_disk, g_filt)
    assert np.array_equal(d.histogram.hbase, g_hbase)
    assert np.array_equal(d.histogram.hbins, g_hbins)
    assert np.array_equal(d.histogram.bound, g_bound)
    assert np.allclose(np.array(d.histogram.peaks), g_peaks, rtol=REL, atol=0)
    assert np.array_equal(np.array(d.histogram.peaks), g_peaks)
    if g_lims is None:
        assert not hasattr(d, "solar_ch_regions")
        return
    keys = list(d.solar_ch_regions.__dict__.keys())
    assert keys == [str(l) for l in g_lims]
    for t, k in enumerate(keys):
        reg = d.solar_ch_regions.__dict__[k]
        assert reg.lim == g_lims[t]
        gp = g_probs[t]
        assert (np.isnan(gp) and np.isnan(reg.prob)) or abs(reg.prob - gp) <= REL * abs(gp), (t, reg.prob, gp)
        assert np.array_equal(reg.raw_map.astype(np.uint8), g_raw[t]), ("raw", t)
        assert np.array_equal(reg.map.astype(np.uint8), g_maps[t]), ("map", t, int((reg.map != g_maps[t]).sum()))
    if g_stack is not None and chips is not None:
        st = chips.probability_map(d)
        assert np.allclose(st, g_stack, rtol=REL, atol=0, equal_nan=True)


@pytest.mark.parametrize("name", ["disk_n256_k11", "disk_n256_k31_211", "disk_n192_k51",
                                  "disk_n256_k5_t28", "disk_n200_k7_wiped", "disk_n192_k11_frange",
                                  "disk_n256_k11_f32", "disk_n192_k31_f32"])
def test_disk_pipeline_vs_reference_golden(name):
    """(the *_f32 goldens: the reference ran on float32 raw / normalized arrays, where numpy keeps
    float32 bin edges, bound, limit_0 and sigmoid; edges, counts, peaks, limits and maps are compared
    exactly, the probabilities to 1e-5 - a float32 exp differs between implementations in the last bit)"""
    g = load_golden(name)
    n = g["image"].shape[0]
    c, d = run_gpu_disk(g["image"], int(g["pixel_radius"]), dtype=in_dtype(g),
                        medfilt_kernel=int(g["kw_medfilt_kernel"]), h_bins=int(g["kw_h_bins"]),
                        threshold_range=[v.item() for v in g["kw_threshold_range"]])
    T = len(g["limits"])
    maps = np.unpackbits(g["maps"])[:T * n * n].reshape(T, n, n)
    raw = np.unpackbits(g["raw_maps"])[:T * n * n].reshape(T, n, n)
    assert np.array_equal(d.solar_mask.i_mask.astype(np.uint8),
                          np.unpackbits(g["i_mask"])[:n * n].reshape(n, n))
    assert np.isnan(d.solar_mask.n_mask[0, 0]) and d.solar_mask.n_mask[n // 2, n // 2] == 1.0
    compare_disk(d, g["filt_disk"].astype(np.float64), g["hbase"], g["hbins"], g["bound"], g["peaks"],
                 g["limits"], g["probs"], raw, maps, g.get("stack"), c)
    assert c.h_thresh == g["h_thresh"]
    assert d.histogram.hbins.dtype == d.histogram.bound.dtype == in_dtype(g)
    assert d.solar_filter.filt_disk.dtype == in_dtype(g)
    reg0 = list(d.solar_ch_regions.__dict__.values())[0]
    assert type(reg0.limit_0) is type(d.histogram.peaks[0]) and reg0.limit_0 == g["limit_0"]
    assert type(reg0.limit_0) is (np.float32 if in_dtype(g) == np.float32 else np.float64)


@pytest.mark.parametrize("n,k,hb,tr,seed", [(512, 11, 500, [0, 20], 7), (384, 31, 1000, [-3, 11], 8),
                                            (320, 51, 500, [0, 20], 9), (512, 3, 20000, [-3, 25], 10),
                                            (256, 7, 500, [-0.75, 9.5], 11), (256, 7, 500, [3, 3], 12)])
def test_disk_pipeline_vs_oracle_seeded(n, k, hb, tr, seed):
    img, r = synth.synthetic_disk(n, seed)
    d_or = synth.FakeDisk(img.astype(np.float64), r)
    co.OracleChips(medfilt_kernel=k, h_bins=hb, threshold_range=tr).run(d_or, keep_contours=False)
    c, d = run_gpu_disk(img, r, medfilt_kernel=k, h_bins=hb, threshold_range=tr)
    if hasattr(d_or, "solar_ch_regions"):
        regs = list(d_or.solar_ch_regions.__dict__.values())
        lims = [x.lim for x in regs]
        probs = [x.prob for x in regs]
        raw = [x.raw_map.astype(np.uint8) for x in regs]
        maps = [x.map.astype(np.uint8) for x in regs]
        st = None
        try:
            st, _ = co.prob_stack([x.map for x in regs], probs)
        except ValueError:
            pass
    else:
        lims = probs = raw = maps = st = None
    compare_disk(d, d_or.solar_filter.filt_disk, d_or.histogram.hbase, d_or.histogram.hbins,
                 d_or.histogram.bound, np.array(d_or.histogram.peaks), lims, probs, raw, maps, st, c)


def test_nan_and_inf_inputs_raise_instead_of_returning_garbage():
    """scipy.signal.medfilt2d is undefined for NaN inside a window (its selection depends on the
    visiting order), so the reference's result cannot be matched: the drop-in classes raise.  A
    NaN that no on-disk window reads (far outside the disk) changes nothing."""
    from pychips_b200 import Chips, SynopticChips
    from pychips_b200._lib import ChipsNativeError
    from pychips_b200 import series
    img, r = synth.synthetic_disk(256, 3)
    c = 128
    for bad, k in ((np.nan, 11), (np.inf, 11), (np.nan, 3), (-np.inf, 1)):
        a = img.astype(np.float64)
        a[c + 5, c - 7] = bad
        d = synth.FakeDisk(a, r)
        with pytest.raises(ChipsNativeError, match="NaN/Inf"):
            Chips(synth.FakeAIA(d), base_folder="/tmp/chips_b200_test/", medfilt_kernel=k).run_CHIPS()
    a = img.astype(np.float64)
    a[0, 0] = np.nan                                     # outside every window the detection reads
    _, d = run_gpu_disk(a, r, medfilt_kernel=11)
    _, d0 = run_gpu_disk(img, r, medfilt_kernel=11)
    assert np.array_equal(d.solar_filter.filt_disk, d0.solar_filter.filt_disk)
    assert list(d.solar_ch_regions.__dict__) == list(d0.solar_ch_regions.__dict__)
    syn = synth.synthetic_synoptic(96, 240, 1).astype(np.float32)
    syn[3, 100] = np.nan                                  # an unfilled polar pixel
    sm = synth.FakeSynoptic(syn)
    with pytest.raises(ChipsNativeError, match="NaN/Inf"):
        SynopticChips(sm, base_folder="/tmp/chips_b200_test/").run_CHIPS()
    run = series.SeriesRunner(256, 256, 11, n_streams=2)
    b = img.copy()
    b[c, c] = np.nan
    with pytest.raises(ChipsNativeError, match="NaN/Inf"):
        run.run([img, b], [r, r])


def test_float64_storage_path_not_f32_representable():
    img, r = synth.synthetic_disk(256, 21)
    img64 = img.astype(np.float64) * (1.0 + 1e-9) + 1e-7       # needs real float64 storage
    assert not np.array_equal(img64.astype(np.float32).astype(np.float64), img64)
    d_or = synth.FakeDisk(img64, r)
    co.OracleChips(medfilt_kernel=11).run(d_or, keep_contours=False)
    c, d = run_gpu_disk(img64, r, medfilt_kernel=11)
    regs = list(d_or.solar_ch_regions.__dict__.values())
    compare_disk(d, d_or.solar_filter.filt_disk, d_or.histogram.hbase, d_or.histogram.hbins,
                 d_or.histogram.bound, np.array(d_or.histogram.peaks), [x.lim for x in regs],
                 [x.prob for x in regs], [x.raw_map.astype(np.uint8) for x in regs],
                 [x.map.astype(np.uint8) for x in regs])


@pytest.mark.parametrize("name", ["syn_180x360_k3", "syn_96x240_k11", "syn_90x200_k5_frange",
                                  "syn_180x360_k3_f32"])
def test_synoptic_pipeline_vs_reference_golden(name):
    from pychips_b200 import SynopticChips
    g = load_golden(name)
    img = g["image"]
    H, W = img.shape
    s = synth.FakeSynoptic(img.astype(in_dtype(g)))
    sc = SynopticChips(s, base_folder="/tmp/chips_b200_test/", medfilt_kernel=int(g["kw_medfilt_kernel"]),
                       h_bins=int(g["kw_h_bins"]),
                       threshold_range=[v.item() for v in g["kw_threshold_range"]])
    sc.run_CHIPS()
    assert np.array_equal(s.solar_filter, g["solar_filter"].astype(np.float64))
    assert np.array_equal(s.histogram.hbase, g["hbase"])
    assert np.array_equal(s.histogram.bound, g["bound"])
    assert np.array_equal(np.array(s.histogram.peaks), g["peaks"])
    keys = list(s.solar_ch_regions.__dict__.keys())
    assert keys == list(g["keys"])
    T = len(keys)
    maps = np.unpackbits(g["maps"])[:T * H * W].reshape(T, H, W)
    for t, k in enumerate(keys):
        reg = s.solar_ch_regions.__dict__[k]
        assert np.array_equal(reg.map.astype(np.uint8), maps[t]), t
        gp = g["probs"][t]
        if in_dtype(g) == np.float32:      # float32 exp: last-bit differences between implementations
            assert (np.isnan(gp) and np.isnan(reg.prob)) or abs(reg.prob - gp) <= REL * abs(gp)
        else:
            assert np.array_equal(reg.prob, gp, equal_nan=True)
    assert s.histogram.bound.dtype == in_dtype(g) and s.solar_filter.dtype == in_dtype(g)


# ----------------------------------------------------------------------------- stage level
@pytest.mark.parametrize("hb", [1, 7, 500, 5000, 20000])
def test_histogram_stage_vs_numpy(hb):
    eng = _engine()
    dev = eng.require_cuda()
    rng = np.random.default_rng(hb)
    x = np.concatenate([rng.normal(28, 3, 30000), rng.normal(130, 15, 70000)]).astype(np.float32)
    img = x.reshape(250, 400)
    det = eng.Detector(eng.Geometry(250, 400, 0, 0, -1), 1, hb, 0, 4, 4, dev)
    t = torch.from_numpy(img).to(dev)
    det.filt.copy_(t)
    det.histogram(recompute_minmax=True)
    det.select_threshold(5, None)
    cn, en = np.histogram(img.astype(np.float64).ravel(), bins=hb)
    hn, _ = np.histogram(img.astype(np.float64).ravel(), bins=hb, density=True)
    assert np.array_equal(det.counts.cpu().numpy(), cn)
    assert np.array_equal(det.edges.cpu().numpy(), en)
    assert np.array_equal(det.density.cpu().numpy(), hn)
    res = det.result()
    pk, _ = signal.find_peaks(hn, height=np.max(hn) / 5)
    assert res.n_peaks == len(pk)
    assert np.array_equal(det.peak_index[:res.n_peaks].cpu().numpy(), pk)
    assert res.h_thresh == np.max(hn) / 5
    assert res.n_samples == img.size


@pytest.mark.parametrize("offset,scale,hb,dtype", [(0.0, 1.0, 500, np.float32), (40.0, 1.0, 500, np.float32),
                                                   (1000.0, 0.004, 500, np.float32), (-300.0, 0.5, 2000, np.float32),
                                                   (1e6, 1.0, 200, np.float64), (3.0, 1e-3, 777, np.float64),
                                                   (0.0, 1.0, 16, np.float32)])
def test_histogram_bin_fast_path_bound_vs_numpy(offset, scale, hb, dtype):
    # k_hist takes trunc((v - first) * nbins / (last - first)) in float32 as the bin when that value is
    # farther than a rounding bound from an integer, and the exact float64 edge comparison otherwise.  The
    # bound grows with |offset| / range: distributions far from zero, narrow ranges, values exactly on bin
    # edges (a uniform grid) and both storage widths must all give numpy's counts exactly.
    eng = _engine()
    dev = eng.require_cuda()
    rng = np.random.default_rng(int(abs(offset)) + hb)
    x = np.concatenate([rng.normal(28, 3, 60000), rng.normal(130, 15, 90000),
                        np.linspace(20.0, 180.0, 10001)])          # the grid puts values ON edges
    x = (offset + scale * x).astype(dtype)
    img = np.ascontiguousarray(x[:160000].reshape(400, 400))
    elem = np.dtype(dtype).itemsize
    det = eng.Detector(eng.Geometry(400, 400, 0, 0, -1), 1, hb, 0, 4, elem, dev)
    det.filt.copy_(torch.from_numpy(img).to(dev))
    det.histogram(recompute_minmax=True)
    det.select_threshold(5, None)
    cn, en = np.histogram(img.astype(np.float64).ravel(), bins=hb)
    assert np.array_equal(det.edges.cpu().numpy(), en)
    got = det.counts.cpu().numpy()
    assert np.array_equal(got, cn), np.flatnonzero(got != cn)[:10]


def test_histogram_constant_image_and_sticky_threshold():
    eng = _engine()
    dev = eng.require_cuda()
    img = np.full((64, 64), 3.0, np.float32)
    det = eng.Detector(eng.Geometry(64, 64, 0, 0, -1), 1, 10, 0, 2, 4, dev)
    det.filt.copy_(torch.from_numpy(img).to(dev))
    det.histogram(recompute_minmax=True)
    det.select_threshold(5, 0.125)
    cn, en = np.histogram(img.astype(np.float64).ravel(), bins=10)
    assert np.array_equal(det.counts.cpu().numpy(), cn)
    assert np.array_equal(det.edges.cpu().numpy(), en)
    assert det.result().h_thresh == 0.125


@pytest.mark.parametrize("seed", list(range(8)) + [100, 101, 102, 103])
def test_cleanup_stage_on_random_levels(seed):
    """Inject a random nested level image and compare W, final maps, labels and 2*areas with
    the first-principles spec (itself pinned to cv2 in test_first_principles.py)."""
    eng = _engine()
    dev = eng.require_cuda()
    from test_first_principles import _random_nested_levels
    rng = np.random.default_rng(1000 + seed)
    n, T = int(rng.integers(40, 300)), int(rng.integers(1, 12))
    r = int(n * rng.uniform(0.3, 0.6))           # > n/2: the disk is clipped by the image border
    c = n // 2
    on = fp.circle_mask(n, n, c, c, r)
    lev = np.where(on, _random_nested_levels(rng, n, T), T).astype(np.uint8)
    thr = float(rng.choice([0.0, 1e-3, 5e-3]))
    if seed >= 100:          # deep trees, "nothing is ever kept" and "everything is kept"
        T = int(rng.integers(20, 40))
        lev = np.where(on, _random_nested_levels(rng, n, T), T).astype(np.uint8)
        thr = [10.0, 0.0, 2e-4, -1.0][seed - 100]
    det = eng.Detector(eng.Geometry(n, n, c, c, r), 3, 8, 0, T, 4, dev)
    det.level.copy_(torch.from_numpy(lev).to(dev))
    det.cleanup(thr)
    # the spec treats everything outside the image as border; off-disk pixels have level T
    W, maps, labels, areas = fp.cleanup_levels(lev, T, r, thr)
    assert np.array_equal(det.fill.cpu().numpy(), W)
    got = det.expand_maps(0).cpu().numpy()
    for t in range(T):
        assert np.array_equal(got[t], maps[t]), (seed, t)
        lab, a2 = det.labels(t)
        assert np.array_equal(lab, labels[t]), (seed, t)
        assert np.array_equal(a2, areas[t][1:]), (seed, t)
    res = det.result()
    for t in range(T):
        assert res.n_comp[t] == len(areas[t]) - 1
        assert res.n_kept[t] == int(((areas[t][1:] * 0.5) / (np.pi * r ** 2) >= thr).sum())


def test_labels_vs_scipy_on_pipeline_output():
    img, r = synth.synthetic_disk(384, 3)
    c, d = run_gpu_disk(img, r, medfilt_kernel=11)
    keys = list(d.solar_ch_regions.__dict__.keys())
    for t in (0, len(keys) // 2, len(keys) - 1):
        reg = d.solar_ch_regions.__dict__[keys[t]]
        lab, a2 = c.labels(d, t)
        ref, n = ndimage.label(reg.filled_map > 0, structure=np.ones((3, 3)))
        assert np.array_equal(lab, ref)
        assert np.array_equal(reg.filled_map > 0, ndimage.binary_fill_holes(reg.raw_map > 0))


# ----------------------------------------------------------------------------- full size
def test_full_size_4096_properties():
    """BASELINE config 2 (4096^2, k=51, h_bins=500, [0,20]): size-independent properties."""
    eng = _engine()
    n = 4096
    img, r = synth.synthetic_disk(n, 0)
    c, d = run_gpu_disk(img, r, dtype=np.float32)
    det = c._detector(d)
    res = det.result()
    assert res.median_errors == 0 and res.n_peaks > 0
    on = fp.circle_mask(n, n, n // 2, n // 2, r)
    assert res.n_samples == int(on.sum())
    filt = det.filt.cpu().numpy()
    # histogram: counts of the GPU == numpy on the GPU-filtered data; first peak -> limit_0.  The
    # arrays are float32 here, so numpy keeps float32 bin edges, bound and limit_0 (bin_type =
    # result_type(a)) and so does the device (plan.f32_math)
    assert filt.dtype == np.float32 and det.f32_math
    cn, en = np.histogram(filt[on], bins=500)
    assert en.dtype == np.float32 and np.array_equal(det.counts.cpu().numpy(), cn)
    assert np.array_equal(det.edges.cpu().numpy(), en.astype(np.float64))
    hn, _ = np.histogram(filt[on], bins=500, density=True)
    assert np.array_equal(det.density.cpu().numpy(), hn)
    pk, _ = signal.find_peaks(hn, height=np.max(hn) / 5)
    l0 = (en[:-1] + np.diff(en))[pk[0]]
    assert l0.dtype == np.float32 and res.limit_0 == l0
    lims = np.array(res.limits[:det.T])
    assert np.array_equal(lims, np.round(l0 + np.arange(0, 20), 4))
    lev = det.level.cpu().numpy()
    assert np.array_equal(lev, fp.level_image(filt, on, lims))
    W = det.fill.cpu().numpy()
    K = det.keep.cpu().numpy()
    assert np.all(W <= lev) and np.all(K >= W)
    assert np.all(K[~on] == det.T)
    maps = det.expand_maps(0).cpu().numpy()
    for t in (0, 5, 19):
        assert np.array_equal(W <= t, ndimage.binary_fill_holes(lev <= t))
        assert np.array_equal(maps[t], (K <= t).astype(np.uint8))
        lab, nl = ndimage.label(W <= t, structure=np.ones((3, 3)))
        a2 = fp.twice_area_quads(W <= t, lab, nl)
        keep = (a2 * 0.5) / (np.pi * r ** 2) >= 1e-3
        keep[0] = False
        assert np.array_equal(maps[t], keep[lab].astype(np.uint8))
    for t in range(1, det.T):
        assert np.all(maps[t] >= maps[t - 1])
    pr = fp.prob_from_counts(fp.prob_counts(filt[on], res.limit_0, lims))
    assert np.allclose(np.array(res.probs[:det.T]), pr, rtol=REL, atol=0, equal_nan=True)


# ----------------------------------------------------------------------------- headline configs vs reference
def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _headline_image(n, g):
    """The fixture stores the sha256 of its input instead of the image (regenerated from the seed)."""
    img, r = synth.synthetic_disk(n, int(g["seed"]))
    assert r == int(g["pixel_radius"])
    if _sha(img) != str(g["image_sha256"]):
        pytest.skip("synth.synthetic_disk is not bit-reproducible on this host (libm/SIMD differ from the "
                    "build container): the reference fixture does not describe this input")
    return img, r


def test_config1_1024_k51_vs_reference_golden():
    """BASELINE config 1 (/root/reference/test/test_method_flow.py:221: k=51, h_bins=500, [0,20]) at
    1024^2 against the outputs of the UNMODIFIED reference (tests/golden/make_golden_headline.py)."""
    g = load_golden("disk_n1024_k51")
    n = 1024
    img, r = _headline_image(n, g)
    c, d = run_gpu_disk(img, r, medfilt_kernel=51, h_bins=500, threshold_range=[0, 20])
    T = len(g["limits"])
    maps = np.unpackbits(g["maps"])[:T * n * n].reshape(T, n, n)
    raw = np.unpackbits(g["raw_maps"])[:T * n * n].reshape(T, n, n)
    compare_disk(d, g["filt_disk"].astype(np.float64), g["hbase"], g["hbins"], g["bound"], g["peaks"],
                 g["limits"], g["probs"], raw, maps)
    assert c.h_thresh == g["h_thresh"]
    assert _sha(c.probability_map(d).astype(np.float64)) == str(g["stack_sha256"])
    regs = list(d.solar_ch_regions.__dict__.values())
    assert [len(x.contours) for x in regs] == list(g["n_contours"])


def test_config2_4096_k51_vs_reference_golden_digest():
    """BASELINE config 2 — the configuration the metric is quoted on (4096^2, reference defaults
    k=51, h_bins=500, [0,20]) — against digests of the UNMODIFIED reference's outputs (about 12
    CPU-minutes in the build container, tests/golden/make_golden_headline.py)."""
    g = load_golden("disk_n4096_k51")
    n = 4096
    img, r = _headline_image(n, g)
    c, d = run_gpu_disk(img, r, medfilt_kernel=51, h_bins=500, threshold_range=[0, 20])
    f32 = d.solar_filter.filt_disk.astype(np.float32)
    assert np.array_equal(f32.astype(np.float64), d.solar_filter.filt_disk)
    if _sha(f32) != str(g["filt_sha256"]):          # localise the mismatch with the 64x64 tile sums
        ts = f32.astype(np.float64).reshape(n // 64, 64, n // 64, 64).sum(axis=(1, 3))
        bad = np.argwhere(ts != g["filt_tile_sums"])
        raise AssertionError(f"filt_disk differs from the reference in {len(bad)} tiles, first {bad[:4].tolist()}")
    assert _sha(np.packbits(d.solar_mask.i_mask.astype(np.uint8))) == str(g["i_mask_sha256"])
    assert np.array_equal(d.histogram.hbase, g["hbase"])
    assert np.array_equal(d.histogram.hbins, g["hbins"])
    assert np.array_equal(d.histogram.bound, g["bound"])
    assert np.array_equal(np.array(d.histogram.peaks), g["peaks"])
    assert c.h_thresh == g["h_thresh"]
    keys = list(d.solar_ch_regions.__dict__.keys())
    assert keys == list(g["keys"])
    for t, k in enumerate(keys):
        reg = d.solar_ch_regions.__dict__[k]
        assert reg.lim == g["limits"][t] and reg.limit_0 == g["limit_0"]
        assert abs(reg.prob - g["probs"][t]) <= REL * abs(g["probs"][t]) and reg.prob == g["probs"][t]
        rm, mp = reg.raw_map.astype(np.uint8), reg.map.astype(np.uint8)
        assert int(rm.sum()) == g["raw_counts"][t] and int(mp.sum()) == g["map_counts"][t], t
        assert _sha(np.packbits(rm)) == str(g["raw_sha256"][t]), ("raw", t)
        assert _sha(np.packbits(mp)) == str(g["map_sha256"][t]), ("map", t)
    assert [len(d.solar_ch_regions.__dict__[k].contours) for k in keys[::7]] == list(g["n_contours"][::7])


# ----------------------------------------------------------------------------- BASELINE configs
def test_config4_synoptic_full_size_vs_oracle():
    """BASELINE config 4: 3600x1440 Carrington map, defaults k=3, h_bins=5000, [-5,20]."""
    from pychips_b200 import SynopticChips
    img = synth.synthetic_synoptic(1440, 3600, 0)
    s_or = synth.FakeSynoptic(img.astype(np.float64))
    co.OracleSynopticChips().run(s_or)
    s = synth.FakeSynoptic(img.astype(np.float64))
    sc = SynopticChips(s, base_folder="/tmp/chips_b200_test/")
    sc.run_CHIPS()
    assert np.array_equal(s.solar_filter, s_or.solar_filter)
    assert np.array_equal(s.histogram.hbase, s_or.histogram.hbase)
    assert np.array_equal(np.array(s.histogram.peaks), np.array(s_or.histogram.peaks))
    assert list(s.solar_ch_regions.__dict__) == list(s_or.solar_ch_regions.__dict__)
    assert len(s.solar_ch_regions.__dict__) == 25
    for k, ref in s_or.solar_ch_regions.__dict__.items():
        got = s.solar_ch_regions.__dict__[k]
        assert np.array_equal(got.map, ref.map), k
        assert np.array_equal(got.prob, ref.prob, equal_nan=True), k
    # the dark region crossing the 0/360 degree seam is detected on both edges
    top = list(s.solar_ch_regions.__dict__.values())[-1].map
    assert top[:, 0].sum() > 0 and top[:, -1].sum() > 0


@pytest.mark.parametrize("k", [11, 31, 51, 71])
def test_config5_sweep_kernel_sizes_16_thresholds(k):
    """BASELINE config 5 (test_ROI_plots.py:72-76 kernels) x 16 thresholds [-3,13]."""
    n = 320
    img, r = synth.synthetic_disk(n, 50 + k)
    d_or = synth.FakeDisk(img.astype(np.float64), r)
    co.OracleChips(medfilt_kernel=k, h_bins=5000, threshold_range=[-3, 13]).run(d_or, keep_contours=False)
    c, d = run_gpu_disk(img, r, medfilt_kernel=k, h_bins=5000, threshold_range=[-3, 13])
    regs = list(d_or.solar_ch_regions.__dict__.values())
    assert len(regs) == 16
    compare_disk(d, d_or.solar_filter.filt_disk, d_or.histogram.hbase, d_or.histogram.hbins,
                 d_or.histogram.bound, np.array(d_or.histogram.peaks), [x.lim for x in regs],
                 [x.prob for x in regs], [x.raw_map.astype(np.uint8) for x in regs],
                 [x.map.astype(np.uint8) for x in regs])


def test_config3_time_series_runner_matches_single_image_path():
    """BASELINE config 3 in miniature: drifting series with jittered radius through
    SeriesRunner (2 streams, H2D per image) == the per-image Chips path; summed probability
    map == sum of the per-image stacks."""
    from pychips_b200 import series
    n, k = 256, 11
    imgs, radii = [], []
    for i in range(5):
        r = int(synth.pixel_radius_for(n) * (1.0 + 0.01 * ((i % 3) - 1)))
        im, _ = synth.synthetic_disk(n, seed=i, radius=r, drift=0.5 * i)
        imgs.append(im)
        radii.append(r)
    runner = series.SeriesRunner(n, n, k, 500, (0, 20), n_streams=2, want_maps=False)
    res = runner.run([torch.from_numpy(im).pin_memory() for im in imgs], radii)
    total = np.zeros((n, n))
    for i, (im, r) in enumerate(zip(imgs, radii)):
        c, d = run_gpu_disk(im, r, medfilt_kernel=k)        # float64 arrays: float64 arithmetic, as the runner's default
        regs = list(d.solar_ch_regions.__dict__.values())
        assert np.array_equal(res.limits[i], [x.lim for x in regs])
        assert np.array_equal(res.probs[i], [x.prob for x in regs], equal_nan=True)
        K = np.full((n, n), len(regs), dtype=np.uint8)
        for t in range(len(regs) - 1, -1, -1):
            K[regs[t].map > 0] = t
        assert np.array_equal(res.keep[i], K)
        st = c.probability_map(d)
        total += np.nan_to_num(st, nan=0.0)
    assert np.allclose(res.prob_map_sum.cpu().numpy(), total, rtol=1e-5, atol=1e-6)
    # the maps above came back sparse (one word per pixel that is in some map, written by the kernel
    # through the pinned mapping); the dense download gives the same arrays
    assert runner.sparse_keep and 0 < res.keep.packed(0).size < n * n // 8
    dense = series.SeriesRunner(n, n, k, 500, (0, 20), n_streams=2, want_maps=False, sparse_keep=False)
    rd = dense.run([torch.from_numpy(im).pin_memory() for im in imgs], radii)
    for i in range(len(imgs)):
        assert np.array_equal(rd.keep[i], res.keep[i])
    # a buffer too small for the words of an image receives the dense map instead
    tiny = series.SeriesRunner(n, n, k, 500, (0, 20), n_streams=2, want_maps=False)
    assert tiny.sparse_cap * 4 >= n * n
    dark = [np.minimum(im, np.float32(np.percentile(im, 30))) for im in imgs[:2]]   # most of the disk is "hole"
    rt = tiny.run([torch.from_numpy(im).pin_memory() for im in dark], radii[:2])
    rdd = dense.run([torch.from_numpy(im).pin_memory() for im in dark], radii[:2])
    assert any(rt.keep.is_dense(i) for i in range(2))
    for i in range(2):
        assert np.array_equal(rt.keep[i], rdd.keep[i])
    # three radii on two stream slots: one workspace per slot, shared by the per-radius plans
    assert len(runner._slots) > 2
    assert len({d.ws.data_ptr() for d in runner._slots.values()}) == 2
    # a second, pipelined pass over the same series (graph replays, submit/result) is identical
    # (result arrays are views of a ring of two pinned buffer sets: copy before submitting twice more)
    res.keep = [k.copy() for k in res.keep]
    h1 = runner.submit([torch.from_numpy(im).pin_memory() for im in imgs], radii)
    h2 = runner.submit([torch.from_numpy(im).pin_memory() for im in imgs[::-1]], radii[::-1])
    r1, r2 = h1.result(), h2.result()
    for i in range(len(imgs)):
        assert np.array_equal(r1.keep[i], res.keep[i]) and np.array_equal(r1.probs[i], res.probs[i], equal_nan=True)
        j = len(imgs) - 1 - i
        assert np.array_equal(r2.keep[j], res.keep[i]) and np.array_equal(r2.limits[j], res.limits[i])


def test_batched_c_abi_entry_matches_the_per_image_path():
    """chips_detect_batch (SURVEY.md §8(b) batched variant; no Python between the images): 7 host
    images with three radii through 3 slots == chips_detect one image at a time (records, level
    maps), the summed probability map == sum of the per-image stacks; also an unmasked handle, a
    float64 handle, and a NaN image reported through its record."""
    import ctypes as C
    from pychips_b200 import _lib
    eng = _engine()
    dev = eng.require_cuda()
    lib = _lib.load()
    n, k, T = 256, 11, 20
    offs = np.arange(0, T).astype(np.float64)

    def run_batch(imgs, cxs, radii, masked, elem, keep=True, psum=True, slots=3):
        h = C.c_void_p()
        H, W = imgs[0].shape
        _lib.check(lib.chips_batch_create(C.byref(h), H, W, k, 500, T, elem, int(masked), slots), "create")
        B = len(imgs)
        pin = [torch.from_numpy(np.ascontiguousarray(im)).pin_memory() for im in imgs]
        ptrs = (C.c_void_p * B)(*[p.data_ptr() for p in pin])
        ci = (C.c_int32 * B)(*[c[0] for c in cxs])
        cj = (C.c_int32 * B)(*[c[1] for c in cxs])
        rr = (C.c_int32 * B)(*radii)
        recs = (_lib.Result * B)()
        keeps = [np.zeros((H, W), np.uint8) for _ in range(B)]
        kp = (C.c_void_p * B)(*[a.ctypes.data for a in keeps]) if keep else None
        ps = np.zeros((H, W), np.float32)
        for _ in range(2):                                   # the second call replays the slot graphs
            rc = lib.chips_detect_batch(h, B, ptrs, 0, ci, cj, rr, 5.0, float("nan"), offs.ctypes.data, 0.8,
                                        1e-3, recs, kp, ps.ctypes.data if psum else None, 0.0)
            _lib.check(rc, "chips_detect_batch")
        lib.chips_batch_destroy(h)
        return recs, keeps, ps

    imgs, radii = [], []
    for i in range(7):
        r = int(synth.pixel_radius_for(n) * (1.0 + 0.02 * ((i % 3) - 1)))
        im, _ = synth.synthetic_disk(n, seed=10 + i, radius=r, drift=0.3 * i)
        imgs.append(im)
        radii.append(r)
    c = n // 2
    recs, keeps, ps = run_batch(imgs, [(c, c)] * 7, radii, True, 4)
    total = np.zeros((n, n))
    for i, (im, r) in enumerate(zip(imgs, radii)):
        det = eng.Detector(eng.Geometry(n, n, c, c, r), k, 500, 0, T, 4, dev)
        det.detect(torch.from_numpy(im).to(dev))
        ref = det.result()
        assert recs[i].median_errors == 0 and recs[i].n_nonfinite == 0 and recs[i].n_peaks == ref.n_peaks > 0
        for f in ("first_edge", "last_edge", "h_thresh", "limit_0", "n_samples"):
            assert getattr(recs[i], f) == getattr(ref, f), (i, f)
        assert list(recs[i].limits[:T]) == list(ref.limits[:T])
        assert np.array_equal(np.array(recs[i].probs[:T]), np.array(ref.probs[:T]), equal_nan=True)
        assert list(recs[i].n_kept[:T]) == list(ref.n_kept[:T])
        assert np.array_equal(keeps[i], det.keep.cpu().numpy())
        total += np.nan_to_num(det.prob_stack(0.0).cpu().numpy(), nan=0.0)
    assert np.allclose(ps, total, rtol=1e-5, atol=1e-6)
    # float64 storage, no outputs but the records
    imgs64 = [im.astype(np.float64) * (1 + 1e-9) for im in imgs[:2]]
    recs64, _, _ = run_batch(imgs64, [(c, c)] * 2, radii[:2], True, 8, keep=False, psum=False, slots=1)
    for i in range(2):
        det = eng.Detector(eng.Geometry(n, n, c, c, radii[i]), k, 500, 0, T, 8, dev)
        det.detect(torch.from_numpy(imgs64[i]).to(dev))
        assert recs64[i].limit_0 == det.result().limit_0 and list(recs64[i].limits[:T]) == list(det.result().limits[:T])
    # unmasked (synoptic) handle
    syn = [synth.synthetic_synoptic(96, 240, s).astype(np.float32) for s in range(2)]
    recs_s, keeps_s, _ = run_batch(syn, [(0, 0)] * 2, [-1, -1], False, 4, psum=False, slots=2)
    for i in range(2):
        det = eng.Detector(eng.Geometry(96, 240, 0, 0, -1), k, 500, 0, T, 4, dev)
        det.detect(torch.from_numpy(syn[i]).to(dev))
        assert recs_s[i].limit_0 == det.result().limit_0
        assert np.array_equal(keeps_s[i], det.keep.cpu().numpy())
    # a NaN is reported in the image's record; a masked handle rejects r < 0
    bad = imgs[0].copy()
    bad[c, c] = np.nan
    recs_b, _, _ = run_batch([imgs[1], bad], [(c, c)] * 2, radii[:2], True, 4, keep=False, psum=False)
    assert recs_b[0].n_nonfinite == 0 and recs_b[1].n_nonfinite > 0
    with pytest.raises(_lib.ChipsNativeError):
        run_batch([imgs[0]], [(c, c)], [-1], True, 4)


def test_h_thresh_is_sticky_across_wavelengths():
    """chips.py:240-241: the first disk's h_thresh is reused for later disks of the same Chips."""
    from pychips_b200 import Chips
    a, r = synth.synthetic_disk(256, 1)
    b, _ = synth.synthetic_disk(256, 2)
    d1 = synth.FakeDisk(a.astype(np.float64), r, wavelength=193)
    d2 = synth.FakeDisk((b * 1.3).astype(np.float32).astype(np.float64), r, wavelength=211)
    c = Chips(synth.FakeAIA([d1, d2]), base_folder="/tmp/chips_b200_test/", medfilt_kernel=11)
    c.run_CHIPS()
    o = co.OracleChips(medfilt_kernel=11)
    e1 = synth.FakeDisk(a.astype(np.float64), r)
    e2 = synth.FakeDisk((b * 1.3).astype(np.float32).astype(np.float64), r)
    o.run(e1, keep_contours=False)
    o.run(e2, keep_contours=False)
    assert d1.histogram.h_thresh == e1.histogram.h_thresh == d2.histogram.h_thresh == e2.histogram.h_thresh
    assert np.array_equal(np.array(d2.histogram.peaks), np.array(e2.histogram.peaks))
    # memoisation: a second run is a no-op unless clear_prev_runs
    before = d1.solar_ch_regions
    c.run_CHIPS(193, 256)
    assert d1.solar_ch_regions is before
    c.run_CHIPS(193, 256, clear_prev_runs=True)
    assert d1.solar_ch_regions is not before
    assert list(d1.solar_ch_regions.__dict__) == list(before.__dict__)


# ----------------------------------------------------------------------------- filament cleanup
def _fl_planes(cf, ret):
    return [np.asarray(a) for a in (ret[0], ret[1], ret[2], ret[3], cf.candidate_bitmap, cf.bmmix,
                                    cf.bmhot, cf.bmcool)]


@pytest.mark.parametrize("n,seed,dtype", [(512, 3, np.float64), (1024, 1, np.float32),
                                          (300, 2, np.float64), (77, 4, np.float64)])
def test_fl_candidates_vs_oracle(n, seed, dtype):
    """CleanFl.create_coronal_hole_candidates (cleanup.py:69-200)."""
    from oracle import cleanfl_oracle as fo
    from pychips_b200.cleanup import CleanFl
    aia = synth.make_triplet_aia(n, seed, dtype=dtype)
    r = aia.datasets[193][n].pixel_radius
    imgs = [np.asarray(aia.datasets[w][n].normalized.data) for w in (171, 193, 211)]
    o = fo.create_coronal_hole_candidates(imgs[0], imgs[1], imgs[2], r, n)
    cf = CleanFl(aia, resolution=n)
    ret = cf.create_coronal_hole_candidates()
    st = cf.stats
    got_thr = np.array([st.thr_mix, st.thr_hot, st.thr_cool])
    got_mean = np.array([st.mean171, st.mean193, st.mean211])
    pow2 = (n * n) % 128 == 0 and ((n * n) // 128) & ((n * n) // 128 - 1) == 0
    if pow2:            # numpy's pairwise summation order reproduced exactly
        assert np.array_equal(got_mean, np.array(o.means)) and np.array_equal(got_thr, np.array(o.thresholds))
    assert np.allclose(got_thr, np.array(o.thresholds), rtol=1e-12, atol=0)
    ref = [o.candidate, o.bmmix, o.bmhot, o.bmcool, o.candidate_bitmap, o.bmmix_m, o.bmhot_m, o.bmcool_m]
    diff = sum(int((a != b).sum()) for a, b in zip(_fl_planes(cf, ret), ref))
    # bit-exact except pixels within one float32 ulp of a threshold (counted by the kernel)
    assert diff <= (0 if pow2 else 8 * st.n_near), (diff, st.n_near)
    assert ret[0].dtype == np.float32 and cf.candidate_bitmap.dtype == np.float64
    assert np.array_equal(ret[4], o.disk_mask)
    on = o.disk_mask > 0
    assert st.n_cand == int(o.candidate[on].sum()) or not pow2
    assert 0.02 < o.candidate[on].mean() < 0.5      # the synthetic case is not degenerate


def test_fl_pipeline_vs_reference_golden():
    """Chips(run_fl_cleanup=True) on the 193 A disk against the unmodified reference."""
    from pychips_b200 import Chips
    g = load_golden("fl_n256_k11")
    n, r = int(g["n"]), int(g["pixel_radius"])
    disks = [synth.FakeDisk(g[k].astype(np.float64), r, w)
             for k, w in (("img171", 171), ("img193", 193), ("img211", 211))]
    aia = synth.FakeAIA(disks)
    c = Chips(aia, base_folder="/tmp/chips_b200_test/", run_fl_cleanup=True,
              medfilt_kernel=int(g["kw_medfilt_kernel"]), h_bins=int(g["kw_h_bins"]),
              threshold_range=[v.item() for v in g["kw_threshold_range"]])
    c.run_CHIPS(wavelength=193, resolution=n)
    planes = np.unpackbits(g["planes"])[:5 * n * n].reshape(5, n, n)
    assert np.array_equal(c.cf.candidate_bitmap.astype(np.uint8), planes[4])
    assert np.array_equal(np.array([c.cf.stats.thr_mix, c.cf.stats.thr_hot, c.cf.stats.thr_cool]),
                          g["thresholds"])
    d = disks[1]
    keys = list(d.solar_ch_regions.__dict__.keys())
    assert keys == list(g["keys"])
    T = len(keys)
    maps = np.unpackbits(g["maps"])[:T * n * n].reshape(T, n, n)
    raw = np.unpackbits(g["raw_maps"])[:T * n * n].reshape(T, n, n)
    for t, k in enumerate(keys):
        reg = d.solar_ch_regions.__dict__[k]
        assert np.array_equal(reg.raw_map.astype(np.uint8), raw[t]), ("raw", t)
        assert np.array_equal(reg.map.astype(np.uint8), maps[t]), ("map", t)
        gp = g["probs"][t]
        assert (np.isnan(gp) and np.isnan(reg.prob)) or abs(reg.prob - gp) <= REL * abs(gp)
    assert maps.sum() > 0 and (maps != raw).any()


def test_fl_requires_three_wavelengths_and_same_size():
    from pychips_b200 import Chips
    from pychips_b200.cleanup import CleanFl
    aia = synth.make_aia(128, 0)
    with pytest.raises(ValueError):
        Chips(aia, base_folder="/tmp/chips_b200_test/", run_fl_cleanup=True).run_CHIPS()
    a3 = synth.make_triplet_aia(128, 0)
    small = synth._Map(np.ones((64, 64)))
    with pytest.raises(NotImplementedError):      # resampling (skimage resize) is out of scope
        CleanFl(a3, resolution=128).create_coronal_hole_candidates(smap171=small)


# ----------------------------------------------------------------------------- outputs (§8(f) rows 2, 4)
def test_region_cube_and_netcdf_file(tmp_path):
    """Chips.to_netcdf (chips.py:512-590): the [res, res, L] f4 cube packed on the device equals
    the reference's `data[:, :, i] = region.map`; the written file round-trips."""
    n = 320
    img, r = synth.synthetic_disk(n, 9)
    d_or = synth.FakeDisk(img.astype(np.float64), r)
    co.OracleChips(medfilt_kernel=11).run(d_or, keep_contours=False)
    want, probs = co.netcdf_arrays(d_or)
    from pychips_b200 import Chips
    d = synth.FakeDisk(img.astype(np.float64), r)
    c = Chips(synth.FakeAIA(d), base_folder=str(tmp_path) + "/", medfilt_kernel=11)
    c.run_CHIPS()
    cube, got_probs = c.region_cube(d)
    assert cube.dtype == np.float32 and cube.shape == want.shape
    assert np.array_equal(cube, want)
    assert np.allclose(got_probs, probs, rtol=REL, atol=0, equal_nan=True)
    path = c.to_netcdf(193, n)
    try:
        from netCDF4 import Dataset
        nc = Dataset(path)
        got = nc["ch_regions"]["ch_regions"][:]
        filt = nc["filter"]["filt_disk"][:]
    except ImportError:
        from scipy.io import netcdf_file
        nc = netcdf_file(path, "r", mmap=False)
        got = nc.variables["ch_regions_ch_regions"][:]
        filt = nc.variables["filter_filt_disk"][:]
        assert nc.variables["fits_fitdata"].shape == (n, n)
        assert np.allclose(nc.variables["ch_regions_probs"][:], np.asarray(probs, dtype=np.float32),
                           equal_nan=True)
    assert np.array_equal(np.asarray(got), want)
    assert np.array_equal(np.asarray(filt), d_or.solar_filter.filt_disk.astype(np.float32))


def test_netcdf4_group_layout_matches_the_reference_call_for_call(tmp_path, monkeypatch):
    """The NETCDF4 branch of Chips.to_netcdf (chips.py:528-588: groups fits / filter / ch_regions,
    their dimensions, variables, dtypes and values) against the call log of the UNMODIFIED
    reference, recorded by tests/golden/make_golden.py with the same stand-in for the absent
    netCDF4 package (oracle/nc_recorder.py)."""
    import json
    import os
    import sys
    from oracle import nc_recorder as ncr
    from pychips_b200 import Chips
    here = os.path.dirname(os.path.abspath(__file__))
    want = json.load(open(os.path.join(here, "golden", "netcdf_layout_n256.json")))
    g = load_golden(want["source"])
    monkeypatch.setitem(sys.modules, "netCDF4", ncr.as_module())
    d = synth.FakeDisk(g["image"].astype(np.float64), int(g["pixel_radius"]))
    c = Chips(synth.FakeAIA(d), base_folder=str(tmp_path) + "/", medfilt_kernel=int(g["kw_medfilt_kernel"]),
              h_bins=int(g["kw_h_bins"]), threshold_range=[v.item() for v in g["kw_threshold_range"]])
    c.run_CHIPS()
    before = len(ncr.LAST)
    c.to_netcdf(193, want["n"])
    assert len(ncr.LAST) == before + 1
    got = json.loads(json.dumps(ncr.LAST[-1].log))
    assert len(got) == len(want["log"])
    for a, b in zip(got, want["log"]):
        assert a == b, (a, b)


@pytest.mark.parametrize("T", [1, 3, 20, 25])
def test_pack_cube_ragged_sizes(T):
    eng = _engine()
    dev = eng.require_cuda()
    H, W = 37, 53
    det = eng.Detector(eng.Geometry(H, W, W // 2, H // 2, 15), 3, 8, 0, T, 4, dev)
    rng = np.random.default_rng(T)
    keep = rng.integers(0, T + 1, size=(H, W)).astype(np.uint8)
    det.keep.copy_(torch.from_numpy(keep).to(dev))
    cube = det.pack_cube(0).cpu().numpy()
    want = (keep[:, :, None] <= np.arange(T)[None, None, :]).astype(np.float32)
    assert np.array_equal(cube, want)


def test_similarity_measures_vs_sklearn():
    """Chips.compute_similarity_measures (chips.py:592-617)."""
    from pychips_b200 import Chips
    rng = np.random.default_rng(3)
    a = (rng.random((200, 333)) < 0.2).astype(np.float64)
    b = (rng.random((200, 333)) < 0.25).astype(np.float64)
    a[5] = 0          # skipped row (sum == 0)
    b[9] = 0          # zero-norm partner -> similarity 0
    a[11] = rng.normal(size=333) + 2.0       # general real-valued rows
    b[11] = rng.normal(size=333) + 2.0
    want = co.compute_similarity_measures(a, b)["cosine_sim"]
    c = Chips(synth.make_aia(64, 0), base_folder="/tmp/chips_b200_test/")
    got = c.compute_similarity_measures(a, b)["cosine_sim"]
    assert abs(got - want) <= 1e-12 * abs(want)
    rows, _, used = _engine().row_cosine(a, b)
    assert used == 199 and np.isnan(rows[5]) and rows[9] == 0.0
    with pytest.warns(RuntimeWarning):
        assert np.isnan(c.compute_similarity_measures(np.zeros((4, 8)), np.ones((4, 8)))["cosine_sim"])


# ----------------------------------------------------------------------------- regional histograms (§8(f) row 3)
@pytest.mark.parametrize("n,xs,ys,hb", [(512, 2, 2, 500), (384, 3, 3, 200), (320, 4, 2, 1000)])
def test_regional_histograms_vs_oracle(n, xs, ys, hb):
    """Chips.extract_histograms (chips.py:260-319, repaired dead code — parity unpinned by the
    reference, pinned to the oracle restatement)."""
    from pychips_b200 import Chips
    img, r = synth.synthetic_disk(n, 12)
    d_or = synth.FakeDisk(img.astype(np.float64), r)
    o = co.OracleChips(medfilt_kernel=11, h_bins=hb)
    o.run(d_or, keep_contours=False)
    regs, fpeak, hthr = co.extract_histograms(d_or.solar_filter.filt_disk, d_or.solar_mask.n_mask, n, xs, ys,
                                              hb, o.h_thresh, 5)
    d = synth.FakeDisk(img.astype(np.float64), r)
    c = Chips(synth.FakeAIA(d), base_folder="/tmp/chips_b200_test/", medfilt_kernel=11, h_bins=hb,
              hist_xsplit=xs, hist_ysplit=ys)
    c.run_CHIPS()
    c.extract_histograms(d)
    assert list(d.histograms.keys) == list(range(xs * ys)) and len(d.histograms.reg) == len(regs)
    assert [int(v) for v in d.histograms.fpeak] == [int(v) for v in fpeak]
    for k, ref in regs.items():
        got = d.histograms.reg[k]
        assert np.array_equal(got.hbase, ref.hbase, equal_nan=True), k
        assert np.array_equal(got.bound, ref.bound), k
        assert np.array_equal(np.asarray(got.peaks), np.asarray(ref.peaks)), k
        assert np.array_equal(got.data, ref.data, equal_nan=True), k
        assert got.h_thresh == ref.h_thresh and got.h_bins == hb
    # the regional pass must not disturb the detection's own record
    d2 = synth.FakeDisk(img.astype(np.float64), r)
    c2 = Chips(synth.FakeAIA(d2), base_folder="/tmp/chips_b200_test/", medfilt_kernel=11, h_bins=hb)
    c2.extract_solar_masks(d2); c2.run_filters(d2); c2.extract_histogram(d2); c2.extract_histograms(d2)
    c2.extract_CHs_CHBs(d2)
    for key, reg in d_or.solar_ch_regions.__dict__.items():
        assert np.array_equal(d2.solar_ch_regions.__dict__[key].map, reg.map)


@pytest.mark.parametrize("n,xs,ys,hb,overlap,sticky", [(512, 2, 2, 500, 0.5, True), (384, 3, 2, 200, 0.75, False),
                                                        (320, 4, 4, 1000, 0.0, True), (256, 1, 1, 5000, 0.5, False)])
def test_sliding_window_histograms_vs_oracle(n, xs, ys, hb, overlap, sticky):
    """Chips.extract_sliding_histograms (an empty stub in the reference, chips.py:321-339 - parity
    unpinned, pinned to the oracle restatement): every window's density, bound, peak indices and
    data, the sticky h_thresh (fixed by the full-disk histogram, or by the first window when the
    method runs on its own), fpeak; all windows in one chips_window_histograms call."""
    from pychips_b200 import Chips
    img, r = synth.synthetic_disk(n, 13)
    d_or = synth.FakeDisk(img.astype(np.float64), r)
    o = co.OracleChips(medfilt_kernel=11, h_bins=hb)
    o.run(d_or, keep_contours=False)
    regs, fpeak, hthr = co.sliding_histograms(d_or.solar_filter.filt_disk, d_or.solar_mask.n_mask, n, xs, ys,
                                              hb, o.h_thresh if sticky else None, 5, overlap)
    d = synth.FakeDisk(img.astype(np.float64), r)
    c = Chips(synth.FakeAIA(d), base_folder="/tmp/chips_b200_test/", medfilt_kernel=11, h_bins=hb,
              hist_xsplit=xs, hist_ysplit=ys)
    if sticky:
        c.run_CHIPS()
    else:
        c.extract_solar_masks(d)
        c.run_filters(d)
    c.extract_sliding_histograms(d, overlap=overlap)
    assert len(regs) == len(co.sliding_windows(n, xs, ys, overlap)) >= 1
    assert list(d.histograms.keys) == list(range(len(regs))) and len(d.histograms.reg) == len(regs)
    assert [int(v) for v in d.histograms.fpeak] == [int(v) for v in fpeak]
    assert c.h_thresh == hthr
    for k, ref in regs.items():
        got = d.histograms.reg[k]
        assert got.window == ref.window
        assert np.array_equal(got.hbase, ref.hbase, equal_nan=True), k
        assert np.array_equal(got.bound, ref.bound), k
        assert np.array_equal(np.asarray(got.peaks), np.asarray(ref.peaks)), k
        assert np.array_equal(got.data, ref.data, equal_nan=True), k
        assert got.h_thresh == ref.h_thresh and got.h_bins == hb


def test_window_histograms_counts_and_empty_windows():
    """chips_window_histograms through the Detector: integer counts of every window == np.histogram,
    a window with no on-disk pixel gives NaN densities and no peaks (numpy on an empty array)."""
    eng = _engine()
    dev = eng.require_cuda()
    n = 384
    img, r = synth.synthetic_disk(n, 4)
    c = n // 2
    det = eng.Detector(eng.Geometry(n, n, c, c, r), 7, 300, 0, 1, 4, dev)
    det.medfilt(torch.from_numpy(img).to(dev))
    filt = det.filt.cpu().numpy().astype(np.float64)
    yy, xx = np.mgrid[:n, :n]
    on = (xx - c) ** 2 + (yy - c) ** 2 <= r * r
    wins = [(0, 40, 0, 40), (100, 300, 50, 384), (0, 384, 0, 384), (191, 193, 10, 380), (300, 384, 300, 384)]
    dens, bnd, pks, hthr, counts = det.window_histograms(wins, 5, None)
    for k, (r0, r1, c0, c1) in enumerate(wins):
        v = filt[r0:r1, c0:c1][on[r0:r1, c0:c1]]
        with np.errstate(invalid="ignore", divide="ignore"):
            h, be = np.histogram(v, bins=300, density=True)
        cnt, _ = np.histogram(v, bins=300)
        assert np.array_equal(counts[k], cnt), k
        assert np.array_equal(dens[k], h, equal_nan=True), k
        assert np.array_equal(bnd[k], be[:-1] + np.diff(be)), k
    assert not on[0:40, 0:40].any() and np.isnan(dens[0]).all() and len(pks[0]) == 0 and np.isnan(hthr)


def _periodic_labels_spec(mask):
    """8-connected components on a cylinder (column W-1 adjacent to column 0), numbered in raster
    order of their first pixel: scipy.ndimage.label + a union of the labels that touch across the seam."""
    H, W = mask.shape
    lab, n = ndimage.label(mask, structure=np.ones((3, 3)))
    parent = list(range(n + 1))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    if W >= 2:
        for y in range(H):
            if not mask[y, 0]:
                continue
            for dy in (-1, 0, 1):
                yy = y + dy
                if 0 <= yy < H and mask[yy, W - 1]:
                    a, b = find(lab[y, 0]), find(lab[yy, W - 1])
                    if a != b:
                        parent[max(a, b)] = min(a, b)
    root = np.array([find(i) for i in range(n + 1)])
    merged = root[lab]
    fg = merged > 0
    flat = merged.ravel()
    order = {}
    for v in flat[fg.ravel()]:
        if v not in order:
            order[v] = len(order) + 1
    lut = np.zeros(n + 1, dtype=np.int32)
    for v, k in order.items():
        lut[v] = k
    out = lut[merged]
    return out, np.bincount(out.ravel(), minlength=len(order) + 1)[1:]


@pytest.mark.parametrize("shape,density,seed", [((96, 240), 0.45, 0), ((64, 64), 0.6, 1), ((33, 2), 0.5, 2),
                                                ((40, 1), 0.5, 3), ((1, 77), 0.5, 4), ((180, 360), 0.3, 5)])
def test_periodic_labels_on_random_masks(shape, density, seed):
    """chips_label_periodic (BASELINE config 4 "wrap-around CCL"; parity unpinned - the reference's
    SynopticChips has no labelling step) against scipy.ndimage.label + a seam merge, and the plain
    (non-periodic) mode against scipy.ndimage.label itself."""
    import ctypes as C
    from pychips_b200 import _lib
    eng = _engine()
    dev = eng.require_cuda()
    lib = _lib.load()
    rng = np.random.default_rng(seed)
    mask = rng.random(shape) < density
    mask[:, 0] |= rng.random(shape[0]) < 0.5          # plenty of contacts across the seam
    mask[:, -1] |= rng.random(shape[0]) < 0.5
    H, W = shape
    img = torch.from_numpy(mask.astype(np.uint8)).to(dev)
    for periodic in (1, 0):
        lab = torch.empty((H, W), dtype=torch.int32, device=dev)
        cap = ((H + 1) // 2) * ((W + 1) // 2)
        areas = torch.empty(cap, dtype=torch.int32, device=dev)
        nlab = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(int(lib.chips_label_workspace_bytes(H, W)), dtype=torch.uint8, device=dev)
        _lib.check(lib.chips_label_periodic(eng._ptr(img), H, W, 1, 0, periodic, eng._ptr(lab), eng._ptr(areas),
                                            cap, eng._ptr(nlab), eng._ptr(ws), eng._stream_ptr()), "label")
        if periodic:
            ref, ref_areas = _periodic_labels_spec(mask)
        else:
            ref, n = ndimage.label(mask, structure=np.ones((3, 3)))
            ref_areas = np.bincount(ref.ravel(), minlength=n + 1)[1:]
        n = int(nlab.item())
        assert n == len(ref_areas)
        assert np.array_equal(lab.cpu().numpy(), ref)
        assert np.array_equal(areas[:n].cpu().numpy(), ref_areas)


def test_synoptic_regions_wrap_around_the_seam():
    """SynopticChips.labels: a dark region that straddles longitude 0 / 360 is ONE region with the
    periodic labelling and two without; labels follow the final map of the chosen threshold."""
    from pychips_b200 import SynopticChips
    syn = synth.synthetic_synoptic(180, 360, 0).astype(np.float32)
    syn[60:90, :14] *= 0.05
    syn[60:90, -11:] *= 0.05
    sm = synth.FakeSynoptic(syn)
    c = SynopticChips(sm, base_folder="/tmp/chips_b200_test/")
    c.run_CHIPS()
    regs = list(sm.solar_ch_regions.__dict__.values())
    t = len(regs) // 2
    mask = regs[t].map > 0
    assert mask[75, 0] and mask[75, -1]
    lab, areas = c.labels(t)
    ref, ref_areas = _periodic_labels_spec(mask)
    assert np.array_equal(lab, ref) and np.array_equal(areas, ref_areas)
    assert lab[75, 0] == lab[75, -1] > 0
    lab0, areas0 = c.labels(t, periodic=False)
    assert lab0[75, 0] != lab0[75, -1] and len(areas0) > len(areas)
    assert int(areas.sum()) == int(mask.sum()) == int(areas0.sum())


# ----------------------------------------------------------------------------- randomised configurations
@pytest.mark.parametrize("seed", range(10))
def test_random_configurations_vs_oracle(seed):
    """Seeded random (size, kernel, bins, thresholds, quantisation, dtype): every stage output of
    the drop-in class against the oracle.  Exercises the register network (k <= 5), the generic
    and the specialised tier-1 / tier-2 kernels and the point buckets inside the full pipeline."""
    rng = np.random.default_rng(4242 + seed)
    n = int(rng.integers(96, 360))
    k = int(rng.choice([3, 5, 7, 9, 11, 21, 31, 51]))
    hb = int(rng.choice([50, 200, 500, 1500]))
    t0 = int(rng.integers(-4, 2))
    tr = [t0, t0 + int(rng.integers(3, 24))]
    q = float(rng.choice([0.0, 0.0, 0.5, 2.0]))
    dtype = np.float64 if rng.random() < 0.5 else np.float32
    img, r = synth.synthetic_disk(n, 300 + seed)
    if q > 0:
        img = (np.round(img / q) * q).astype(np.float32)
    if dtype == np.float64 and rng.random() < 0.5:
        img = img.astype(np.float64) * (1.0 + 2.0 ** -40)      # not float32-representable
    d_or = synth.FakeDisk(img.astype(np.float64), r)
    co.OracleChips(medfilt_kernel=k, h_bins=hb, threshold_range=tr).run(d_or, keep_contours=False)
    c, d = run_gpu_disk(img.astype(np.float64), r, medfilt_kernel=k, h_bins=hb, threshold_range=tr)
    if hasattr(d_or, "solar_ch_regions"):
        regs = list(d_or.solar_ch_regions.__dict__.values())
        lims, probs = [x.lim for x in regs], [x.prob for x in regs]
        raw = [x.raw_map.astype(np.uint8) for x in regs]
        maps = [x.map.astype(np.uint8) for x in regs]
    else:
        lims = probs = raw = maps = None
    compare_disk(d, d_or.solar_filter.filt_disk, d_or.histogram.hbase, d_or.histogram.hbins,
                 d_or.histogram.bound, np.array(d_or.histogram.peaks), lims, probs, raw, maps, None, c)


def test_graph_replay_matches_plain_launches():
    """Detector.detect_graphed: the captured launch sequence, replayed on NEW contents of the same
    buffers, gives exactly what the plain launches give (records, maps, launch accounting)."""
    eng = _engine()
    dev = eng.require_cuda()
    n = 320
    imgs = [synth.synthetic_disk(n, 40 + i) for i in range(3)]
    r = imgs[0][1]
    buf = torch.empty((n, n), dtype=torch.float32, device=dev)
    maps = torch.empty((20, n, n), dtype=torch.uint8, device=dev)
    det = eng.Detector(eng.Geometry(n, n, n // 2, n // 2, r), 11, 500, 0, 20, 4, dev)
    ref = eng.Detector(eng.Geometry(n, n, n // 2, n // 2, r), 11, 500, 0, 20, 4, dev)
    rmaps = torch.empty_like(maps)
    for i, (img, _) in enumerate(imgs):
        buf.copy_(torch.from_numpy(img))
        det.detect_graphed(buf, maps=maps)        # first call: plain + capture, then replays
        ref.detect(buf, maps=rmaps)
        torch.cuda.synchronize()
        assert torch.equal(det.keep, ref.keep) and torch.equal(maps, rmaps), i
        a, b = det.result(), ref.result()           # (entries beyond T are never written)
        for f in ("first_edge", "last_edge", "h_thresh", "limit_0", "n_samples", "n_peaks", "median_errors"):
            assert getattr(a, f) == getattr(b, f), (i, f)
        for f in ("limits", "probs", "n_kept", "n_comp"):
            assert np.array_equal(np.array(getattr(a, f)[:20]), np.array(getattr(b, f)[:20]), equal_nan=True), (i, f)
    assert det.graph_launches > 0 and len(det._graphs) == 1


# ----------------------------------------------------------------------------- contours (§8(f) row 4)
def _same_contours(mask, got, hier):
    import cv2
    ref, hi = cv2.findContours(mask, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    assert len(ref) == len(got)
    for k, (a, b) in enumerate(zip(ref, got)):
        assert a.dtype == b.dtype and np.array_equal(a, b), k
    if len(ref):
        assert np.array_equal(hi, hier)
    else:
        assert hier is None
    return ref


@pytest.mark.parametrize("seed", range(4))
def test_contours_random_masks_vs_opencv(seed):
    """chips_contours (csrc/contours.cu) == cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE): points,
    order and hierarchy, on noise (thousands of borders, 1-pixel walls), blobs and ragged shapes."""
    eng = _engine()
    rng = np.random.default_rng(300 + seed)
    shapes = [(1, 1), (1, 17), (23, 1), (37, 53), (64, 64), (130, 257), (200, 333)]
    for h, w in shapes:
        for p in (0.25, 0.5, 0.75):
            m = (rng.random((h, w)) < p).astype(np.uint8)
            _same_contours(m, *eng.find_contours(m))
    yy, xx = np.mgrid[:300, :410]
    blobs = ((np.sin(xx / (7.0 + seed)) * np.cos(yy / 5.0) > 0.2) |
             ((xx - 200) ** 2 + (yy - 150) ** 2 < 90 ** 2)).astype(np.uint8)
    blobs[120:180, 170:230] = 0
    blobs[140:160, 190:210] = 255
    assert len(_same_contours(blobs, *eng.find_contours(blobs))) > 10
    for m in (np.zeros((9, 12), np.uint8), np.ones((9, 12), np.uint8)):
        _same_contours(m, *eng.find_contours(m))


def test_contours_capacities_grow_on_demand():
    eng = _engine()
    dev = eng.require_cuda()
    from pychips_b200 import contours as pc
    rng = np.random.default_rng(7)
    m = (rng.random((150, 170)) < 0.55).astype(np.uint8)
    tr = eng.ContourTracer(170, 150, dev, max_border_px=64, max_contours=8, max_points=16)
    tr.start_rounds = 2
    recs, pts = tr.trace(torch.from_numpy(m).to(dev), 1, 0)
    got, hier, area2 = pc.assemble(recs, pts)
    ref = _same_contours(m, got, hier[None])
    assert len(ref) > 1000 and tr.calls >= 3
    import cv2
    assert np.array_equal(area2, [int(round(2 * cv2.contourArea(c))) for c in ref])
    with pytest.raises(ValueError):
        tr.trace(torch.zeros((10, 10), dtype=torch.uint8, device=dev))


def test_contours_lattice_needs_many_start_rounds():
    """A lattice of 1-pixel walls: the start positions settle only after O(image side) rounds; the
    tracer re-enqueues with more rounds until the call reports that nothing moved."""
    eng = _engine()
    m = np.zeros((200, 230), np.uint8)
    m[::2, :] = 1
    m[:, ::3] = 1
    dev = eng.require_cuda()
    tr = eng.ContourTracer(230, 200, dev)
    from pychips_b200 import contours as pc
    recs, pts = tr.trace(torch.from_numpy(m).to(dev), 1, 0)
    got, hier, _ = pc.assemble(recs, pts)
    assert len(_same_contours(m, got, hier[None])) > 1000 and tr.calls >= 4


@pytest.mark.parametrize("n,k,seed", [(320, 11, 3), (512, 31, 8)])
def test_contours_of_the_pipeline_vs_oracle(n, k, seed, tmp_path):
    """region.contours / hierarchy / contour_ids of Chips.extract_CHs_CHBs (chips.py:382-400) against
    the reference's cv2 calls (oracle with keep_contours=True), every threshold."""
    img, r = synth.synthetic_disk(n, seed)
    d_or = synth.FakeDisk(img.astype(np.float64), r)
    co.OracleChips(medfilt_kernel=k).run(d_or, keep_contours=True)
    from pychips_b200 import Chips
    d = synth.FakeDisk(img.astype(np.float64), r)
    Chips(synth.FakeAIA(d), base_folder=str(tmp_path) + "/", medfilt_kernel=k).run_CHIPS()
    want, got = vars(d_or.solar_ch_regions), vars(d.solar_ch_regions)
    assert list(want) == list(got) and len(want) == 20
    total = 0
    for key in want:
        a, b = want[key], got[key]
        assert len(a.contours) == len(b.contours), key
        for x, y in zip(a.contours, b.contours):
            assert x.dtype == y.dtype and np.array_equal(x, y), key
        assert np.array_equal(np.asarray(a.hierarchy), np.asarray(b.hierarchy)), key
        assert a.contour_ids == b.contour_ids
        total += len(a.contours)
    assert total > 0


def test_contours_full_size_vs_opencv():
    """4096^2 detection: every border of two threshold masks against OpenCV."""
    import cv2
    eng = _engine()
    dev = eng.require_cuda()
    n = 4096
    img, r = synth.synthetic_disk(n, 1)
    det = eng.Detector(eng.Geometry(n, n, n // 2, n // 2, r), 51, 500, 0, 20, 4, dev)
    det.detect(torch.from_numpy(img).to(dev))
    torch.cuda.synchronize()
    lev = det.level.cpu().numpy()
    for t in (0, 19):
        mask = (lev <= t).astype(np.uint8)
        got, hier, area2 = det.contours(t)
        ref = _same_contours(mask, got, hier[None] if len(got) else None)
        assert np.array_equal(area2, [int(round(2 * cv2.contourArea(c))) for c in ref])
        assert len(ref) > 0
    # all thresholds in one enqueue == one at a time
    every = det.contours_all()
    assert len(every) == 20
    for t in (0, 7, 19):
        got, hier, area2 = det.contours(t)
        g2, h2, a2 = every[t]
        assert len(got) == len(g2) and all(np.array_equal(a, b) for a, b in zip(got, g2))
        assert np.array_equal(hier, h2) and np.array_equal(area2, a2)


# ----------------------------------------------------------------------------- uploads
@pytest.mark.parametrize("n,k,rfrac", [(256, 11, 0.385), (320, 31, 0.25), (200, 5, 0.3), (256, 11, 0.6)])
def test_series_uploads_only_the_rectangle_the_kernels_read(n, k, rfrac):
    """SeriesRunner.roi_copy: only Detector.input_rect of each host image is uploaded.  With the
    device input buffers poisoned (NaN) beforehand, records and maps equal the full-copy path —
    i.e. nothing outside the rectangle is ever read."""
    from pychips_b200 import series
    r = int(n * rfrac)
    imgs = [synth.synthetic_disk(n, seed=70 + i, radius=r)[0] for i in range(3)]
    pinned = [torch.from_numpy(im).pin_memory() for im in imgs]
    out = []
    for roi in (False, True):
        runner = series.SeriesRunner(n, n, k, 500, (0, 20), n_streams=2, want_maps=False)
        runner.roi_copy = roi
        runner._ensure_copy_pipeline()
        for pair in runner._dev_in2:
            for buf in pair:
                buf.fill_(float("nan"))
        torch.cuda.synchronize()
        res = runner.run(pinned, [r] * 3)
        det0 = next(iter(runner._slots.values()))
        out.append((res, runner.h2d_bytes, det0.input_rect, det0.input_spans))
    (a, bytes_full, _, _), (b, bytes_roi, rect, spans) = out
    x0, y0, x1, y1 = rect
    assert bytes_full == 3 * n * n * 4
    assert bytes_roi == 3 * 4 * sum((sy1 - sy0) * (sx1 - sx0) for (sy0, sy1, sx0, sx1) in spans)
    assert bytes_roi <= 3 * (x1 - x0) * (y1 - y0) * 4
    assert bytes_roi < bytes_full or r > n // 2
    assert int(a.n_peaks.min()) > 0
    for i in range(3):
        assert np.array_equal(a.keep[i], b.keep[i])
        assert np.array_equal(a.limits[i], b.limits[i]) and np.array_equal(a.probs[i], b.probs[i], equal_nan=True)
    assert torch.equal(a.prob_map_sum, b.prob_map_sum)
