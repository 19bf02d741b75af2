"""CPU: host-side logic — the drop-in classes fail loudly without a GPU (no fallback), the
lazy result namespace, synthetic generators, series sharding and the 2-rank gloo exchange."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from pychips_b200 import synth


def test_constructor_signatures_match_reference_defaults():
    import inspect
    from pychips_b200 import Chips, SynopticChips
    sig = inspect.signature(Chips.__init__).parameters
    ref = dict(base_folder="tmp/chips_dataset/", medfilt_kernel=51, h_bins=500, h_thresh=None,
               ht_peak_ratio=5, hist_xsplit=2, hist_ysplit=2, threshold_range=[0, 20],
               porb_threshold=0.8, area_threshold=1e-3, run_fl_cleanup=False)
    assert list(sig)[:2] == ["self", "aia"]
    for k, v in ref.items():
        assert sig[k].default == v, k
    assert list(sig)[2:2 + len(ref)] == list(ref)          # same positional order (chips.py:47-61)
    sig = inspect.signature(SynopticChips.__init__).parameters
    ref = dict(base_folder="tmp/chips_dataset/", medfilt_kernel=3, h_bins=5000, h_thresh=None,
               ht_peak_ratio=5, hist_xsplit=2, hist_ysplit=2, threshold_range=[-5, 20],
               porb_threshold=0.8)
    for k, v in ref.items():
        assert sig[k].default == v, k
    assert list(sig)[2:2 + len(ref)] == list(ref)          # syn_chips.py:40-52
    for name in ("run_CHIPS", "process_CHIPS", "extract_solar_masks", "run_filters",
                 "extract_histogram", "extract_CHs_CHBs", "clear_run_results", "get_base_folder"):
        assert callable(getattr(Chips, name))
    for name in ("run_CHIPS", "run_filters", "extract_histogram", "extract_CHs_CHBs"):
        assert callable(getattr(SynopticChips, name))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu(tmp_path):
    from pychips_b200 import Chips, SynopticChips, engine
    from pychips_b200._lib import ChipsNativeError
    img, r = synth.synthetic_disk(64, 0)
    d = synth.FakeDisk(img.astype(np.float64), r)
    with pytest.raises(ChipsNativeError):
        Chips(synth.FakeAIA(d), base_folder=str(tmp_path) + "/")
    with pytest.raises(ChipsNativeError):
        SynopticChips(synth.FakeSynoptic(img), base_folder=str(tmp_path) + "/")
    with pytest.raises(ChipsNativeError):
        engine.medfilt2d(img, 3)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pychips_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.ChipsNativeError):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pychips_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                # no CPU numerics library (scipy / OpenCV) on the product path
                assert not re.search(r"^\s*(import|from)\s+(scipy|cv2)\b", src, flags=re.M), f


def test_lazy_namespace():
    from pychips_b200.chips import LazyNamespace
    calls = []
    ns = LazyNamespace(_lazy=dict(big=lambda: calls.append(1) or np.arange(3)), lim=2.5)
    assert ns.lim == 2.5 and "big" in ns and "lim" in ns and "nope" not in ns
    assert calls == []
    assert np.array_equal(ns.big, np.arange(3)) and np.array_equal(ns.big, np.arange(3))
    assert calls == [1]
    with pytest.raises(AttributeError):
        ns.nope
    assert hasattr(ns, "big") and not hasattr(ns, "nope")


def test_synthetic_inputs_are_seeded_and_f32():
    a, r = synth.synthetic_disk(128, 3)
    b, _ = synth.synthetic_disk(128, 3)
    c, _ = synth.synthetic_disk(128, 4)
    assert a.dtype == np.float32 and np.array_equal(a, b) and not np.array_equal(a, c)
    assert r == int(0.385 * 128) and np.isfinite(a).all()
    s = synth.synthetic_synoptic(90, 180, 1)
    assert s.shape == (90, 180) and s.dtype == np.float32 and np.isfinite(s).all()
    assert (s[:, 0] < 40).any() and (s[:, -1] < 40).any()       # dark region across the seam
    aia = synth.make_aia(64, 0)
    assert aia.wavelengths == [193] and aia.datasets[193][64].resolution == 64


def test_shard_indices_partition():
    from pychips_b200.series import shard_indices
    for n, w in [(512, 8), (10, 4), (3, 8), (0, 2)]:
        parts = [shard_indices(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _gloo_worker(rank, world, port, n_total, T, H, W, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pychips_b200 import series
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = series.shard_indices(n_total, rank, world)
    n = len(idx)
    rng = np.random.default_rng(rank)
    res = series.SeriesResult(
        indices=idx, limit_0=np.array([20.0 + i for i in idx]), h_thresh=np.zeros(n),
        n_peaks=np.ones(n), limits=np.array([[i + t for t in range(T)] for i in idx], dtype=float).reshape(n, T),
        probs=np.array([[0.01 * i * t for t in range(T)] for i in idx]).reshape(n, T),
        n_kept=np.array([[i % 3 + t for t in range(T)] for i in idx]).reshape(n, T))
    local = torch.full((H, W), float(rank + 1))
    allstats, psum = series.gather_series(series.pack_stats(res, T), local, n_total, T)
    # a second batch's exchange with the same per-rank buffer: the first reduction must not have
    # been folded into it (the round-1 code all-reduced in place: 3, then 6)
    _, psum2 = series.gather_series(series.pack_stats(res, T), local, n_total, T)
    q.put((rank, allstats, psum.numpy(), psum2.numpy(), local.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_series_exchange_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_total, T, H, W = 7, 4, 6, 5
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_total, T, H, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, allstats, psum, psum2, local in outs:
        assert np.all(local == rank + 1.0)                                  # this rank's map is untouched
        assert np.all(psum2 == 3.0)                                         # two batches: 3 and 3, not 3 and 6
        assert allstats.shape == (n_total, 3 + 3 * T)
        assert np.array_equal(allstats[:, 0], np.arange(n_total))          # ordered by image index
        assert np.array_equal(allstats[:, 1], 20.0 + np.arange(n_total))
        assert np.array_equal(allstats[:, 3:3 + T], np.arange(n_total)[:, None] + np.arange(T)[None])
        assert np.all(psum == 3.0)                                          # 1 + 2 summed over ranks


def test_netcdf3_writer_round_trip(tmp_path):
    """pychips_b200/ncio.py (used by Chips.to_netcdf when netCDF4 is absent) against scipy's reader."""
    from scipy.io import netcdf_file
    from pychips_b200 import ncio
    rng = np.random.default_rng(0)
    a = rng.random((5, 7)).astype(np.float32)
    c = rng.random((5, 7, 3))
    p = np.array([0.1, np.nan, 0.3])
    path = str(tmp_path / "t.nc")
    ncio.write_float32(path, dict(xarcs=5, yarcs=7, probs=3),
                       {"fits_fitdata": (("xarcs", "yarcs"), a),
                        "ch_regions_ch_regions": (("xarcs", "yarcs", "probs"), c),
                        "ch_regions_probs": (("probs",), p)})
    nc = netcdf_file(path, "r", mmap=False)
    assert nc.dimensions == {"xarcs": 5, "yarcs": 7, "probs": 3}
    assert np.array_equal(nc.variables["fits_fitdata"][:], a)
    assert np.array_equal(nc.variables["ch_regions_ch_regions"][:], c.astype(np.float32))
    assert np.array_equal(nc.variables["ch_regions_probs"][:], p.astype(np.float32), equal_nan=True)
    with pytest.raises(ValueError):
        ncio.write_float32(path, dict(x=2), {"v": (("x",), np.zeros(3))})
