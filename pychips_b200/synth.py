"""Seeded synthetic inputs for the CHIPS hot path.

Everything is generated with ``np.random.default_rng(seed)`` **as float32** so that
the B200 path (fp32 storage, fp64 compares in registers) and the CPU oracle
(float64 arrays, as the reference receives from aiapy) see bit-identical values:
widen with ``img.astype(np.float64)`` for the oracle.

* ``synthetic_disk``      – AIA-193-like limb-brightened full disk with injected dark
  regions (one with a bright island, one touching the limb).  The container the
  reference builds from FITS (`/root/reference/chips/fetch.py:38-78,194-214`) is replaced
  by ``FakeDisk`` / ``FakeAIA`` duck types below (contract: SURVEY.md §8(b)).
* ``synthetic_synoptic``  – Carrington synoptic map [1440, 3600] with a dark region that
  crosses the 0/360 degree seam (`/root/reference/chips/fetch.py:433-499` replaced by
  ``FakeSynoptic``).
"""
from __future__ import annotations

import datetime as dt
from argparse import Namespace

import numpy as np


def pixel_radius_for(n: int) -> int:
    """`pixel_radius = int(rsun_obs / cdelt2)` in the reference (fetch.py:194-214);
    0.385*N is the value the reference's own scripts use (test/test_proposal.py:74)."""
    return int(0.385 * n)


def _ellipse(xx, yy, x0, y0, a, b, th):
    c, s = np.cos(th), np.sin(th)
    u = (xx - x0) * c + (yy - y0) * s
    v = -(xx - x0) * s + (yy - y0) * c
    return (u / a) ** 2 + (v / b) ** 2


def _ellipse_mask(xx, yy, x0, y0, a, b, th):
    """`_ellipse(...) <= 1` evaluated only inside the ellipse's bounding box (same values)."""
    n0, n1 = xx.shape
    rad = max(a, b) + 2
    i0, i1 = max(0, int(np.floor(y0 - rad))), min(n0, int(np.ceil(y0 + rad)) + 1)
    j0, j1 = max(0, int(np.floor(x0 - rad))), min(n1, int(np.ceil(x0 + rad)) + 1)
    m = np.zeros(xx.shape, dtype=bool)
    if i0 < i1 and j0 < j1:
        m[i0:i1, j0:j1] = _ellipse(xx[i0:i1, j0:j1], yy[i0:i1, j0:j1], x0, y0, a, b, th) <= 1.0
    return m


def synthetic_disk(n: int = 1024, seed: int = 0, radius: int | None = None,
                   drift: float = 0.0) -> tuple[np.ndarray, int]:
    """Return (image float32 [n, n], pixel_radius).

    ``drift`` moves the dark-region geometry slowly (time-series configs)."""
    rng = np.random.default_rng(seed)
    r = pixel_radius_for(n) if radius is None else int(radius)
    cx = cy = int(n / 2)
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32)
    rho = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2) / np.float32(r)
    on = rho <= 1.0
    mu = np.sqrt(np.clip(1.0 - rho * rho, 0.0, 1.0))
    img = np.where(on, 120.0 * (1.0 + 0.6 * (1.0 - mu)),
                   8.0 * np.exp(-6.0 * np.clip(rho - 1.0, 0.0, None))).astype(np.float32)
    # geometry is drawn from a generator that does NOT depend on the image index when
    # drifting, so a series shares its regions and only moves them.
    grng = np.random.default_rng(12345 if drift else seed + 1000003)
    nreg = int(grng.integers(3, 7))
    dark = np.zeros((n, n), dtype=bool)
    for i in range(nreg):
        ang = grng.uniform(0, 2 * np.pi)
        if i == 1:                      # one region touches the limb
            rad = 0.93 * r
        else:
            rad = grng.uniform(0.05, 0.7) * r
        x0 = cx + rad * np.cos(ang) + drift * (0.02 * r) * (i + 1)
        y0 = cy + rad * np.sin(ang)
        a = grng.uniform(0.10, 0.30) * r
        b = grng.uniform(0.07, 0.20) * r
        th = grng.uniform(0, np.pi) + 0.01 * drift
        reg = _ellipse_mask(xx, yy, x0, y0, a, b, th)
        if i == 0:                      # bright island inside the first region
            reg &= ~_ellipse_mask(xx, yy, x0, y0, 0.35 * a, 0.35 * b, th)
            # and a dark core inside the island (nested hole/island path)
            reg |= _ellipse_mask(xx, yy, x0, y0, 0.12 * a, 0.12 * b, th)
        dark |= reg
    img = np.where(dark & on, np.float32(25.0), img).astype(np.float32)
    img += rng.gamma(2.0, 2.0, size=(n, n)).astype(np.float32)
    return np.ascontiguousarray(img, dtype=np.float32), r


def synthetic_synoptic(h: int = 1440, w: int = 3600, seed: int = 0) -> np.ndarray:
    """Carrington synoptic map, float32 [h, w], NaN-free, with one dark region across
    the longitude seam."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    lat = (yy / h - 0.5) * np.pi
    lon = xx / w * 2 * np.pi
    img = (110.0 + 12.0 * np.cos(2 * lat) + 6.0 * np.sin(3 * lon) * np.cos(lat)).astype(np.float32)
    grng = np.random.default_rng(seed + 7777)
    dark = np.zeros((h, w), dtype=bool)
    for i in range(int(grng.integers(4, 8))):
        x0 = 0.0 if i == 0 else grng.uniform(0, w)
        y0 = grng.uniform(0.15, 0.85) * h
        a = grng.uniform(0.02, 0.06) * w
        b = grng.uniform(0.03, 0.10) * h
        th = grng.uniform(0, np.pi)
        for shift in (-w, 0, w):       # periodic in longitude
            dark |= _ellipse(xx, yy, x0 + shift, y0, a, b, th) <= 1.0
    img = np.where(dark, np.float32(22.0), img).astype(np.float32)
    img += rng.gamma(2.0, 2.0, size=(h, w)).astype(np.float32)
    return np.ascontiguousarray(img, dtype=np.float32)


class _Map:
    """Stand-in for `sunpy.map.Map`: only `.data` is read on the hot path
    (`/root/reference/chips/chips.py:184,213-215`)."""

    def __init__(self, data):
        self.data = data


class FakeDisk:
    """Duck type of `chips.fetch.SolarDisk` (`/root/reference/chips/fetch.py:38-91`):
    `.raw.data`, `.normalized.data`, `.resolution`, `.pixel_radius`, `.wavelength`,
    `.date`, `set_value/get_value`."""

    def __init__(self, data, pixel_radius, wavelength=193, date=None):
        self.raw = _Map(data)
        self.normalized = _Map(data)
        self.resolution = int(data.shape[0])
        self.pixel_radius = int(pixel_radius)
        self.wavelength = wavelength
        self.date = date or dt.datetime(2018, 5, 30, 12)

    def set_value(self, key, value):
        setattr(self, key, value)

    def get_value(self, key):
        return getattr(self, key)


class FakeAIA:
    """Duck type of `chips.fetch.RegisterAIA` (`/root/reference/chips/fetch.py:274-316`)."""

    def __init__(self, disks, date=None):
        disks = disks if isinstance(disks, (list, tuple)) else [disks]
        self.date = date or disks[0].date
        self.resolution = disks[0].resolution
        self.wavelengths = [d.wavelength for d in disks]
        self.datasets = {d.wavelength: {d.resolution: d} for d in disks}


class FakeSynoptic:
    """Duck type of `chips.fetch.SynopticMap` (`/root/reference/chips/fetch.py:433-523`)."""

    def __init__(self, data, wavelength=193, date=None):
        self.raw = _Map(data)
        self.wavelength = wavelength
        self.date = date or dt.datetime(2018, 5, 30, 12)

    def set_value(self, key, value):
        setattr(self, key, value)

    def get_value(self, key):
        return getattr(self, key)


def make_aia(n=1024, seed=0, dtype=np.float64, wavelength=193):
    img, r = synthetic_disk(n, seed)
    return FakeAIA(FakeDisk(img.astype(dtype), r, wavelength))


def synthetic_triplet(n: int = 1024, seed: int = 0):
    """Three co-registered AIA-like images (171, 193, 211 A) of the same synthetic Sun, float32,
    for the filament-cleanup path (`/root/reference/chips/cleanup.py:69-200`): coronal holes are
    dark in all three bands, "filaments" are dark in 193 but keep a cool (171-bright) signature,
    so the CHIMERA ratios separate them.  Returns ({171: img, 193: img, 211: img}, pixel_radius)."""
    i193, r = synthetic_disk(n, seed)
    rng = np.random.default_rng(seed + 424242)
    cx = cy = n // 2
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32)
    on = (xx - cx) ** 2 + (yy - cy) ** 2 <= np.float32(r) ** 2
    base = np.maximum(i193 - rng.gamma(2.0, 2.0, size=(n, n)).astype(np.float32) * 0.5, 0.5)
    # filament-like strips: dark in 193/211, nearly unchanged in 171
    fil = np.zeros((n, n), dtype=bool)
    frng = np.random.default_rng(seed + 99)
    for _ in range(int(frng.integers(2, 5))):
        ang = frng.uniform(0, 2 * np.pi)
        rad = frng.uniform(0.1, 0.8) * r
        fil |= _ellipse_mask(xx, yy, cx + rad * np.cos(ang), cy + rad * np.sin(ang),
                             frng.uniform(0.08, 0.2) * r, frng.uniform(0.01, 0.03) * r,
                             frng.uniform(0, np.pi))
    fil &= on
    i193 = np.where(fil, np.float32(30.0) + i193 * np.float32(0.05), i193).astype(np.float32)
    i171 = (1.9 * base ** 0.92 + rng.gamma(2.0, 3.0, size=(n, n))).astype(np.float32)
    i211 = (0.02 * base ** 1.7 + rng.gamma(2.0, 0.8, size=(n, n))).astype(np.float32)
    i211 = np.where(fil, i211 * np.float32(0.35), i211).astype(np.float32)
    # a few non-positive pixels exercise clip_negative / log10(0) = -inf (cleanup.py:131-139)
    bad = rng.random((n, n)) < 2e-4
    i171 = np.where(bad, np.float32(-1.0), i171).astype(np.float32)
    i211 = np.where(np.roll(bad, 3, axis=1), np.float32(0.0), i211).astype(np.float32)
    return {171: np.ascontiguousarray(i171), 193: np.ascontiguousarray(i193),
            211: np.ascontiguousarray(i211)}, r


def make_triplet_aia(n=1024, seed=0, dtype=np.float64):
    imgs, r = synthetic_triplet(n, seed)
    return FakeAIA([FakeDisk(imgs[w].astype(dtype), r, w) for w in (171, 193, 211)])
