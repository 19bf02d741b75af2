"""TEST INFRASTRUCTURE — not product code.

Imports the UNMODIFIED reference (`/root/reference/chips/chips.py`,
`/root/reference/chips/syn_chips.py`) with `MagicMock` stand-ins for the packages that
only its non-hot-path code needs, and runs the stage methods on a
synthetic container.  Only usable where `/root/reference` exists (this container);
`tests/golden/make_golden.py` uses it to pin `oracle/chips_oracle.py` and to write the
committed golden fixtures.  Nothing under `pychips_b200/` may import this.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

REF = "/root/reference"
_STUBS = [
    "aiapy", "aiapy.psf", "aiapy.calibrate", "astropy", "astropy.units", "astropy.io",
    "astropy.io.fits", "sunpy", "sunpy.io", "sunpy.map", "sunpy.coordinates",
    "sunpy.coordinates.sun", "sunpy.net", "sunpy.visualization",
    "sunpy.visualization.colormaps", "matplotlib", "matplotlib.pyplot", "matplotlib.axes",
    "matplotlib.image", "matplotlib.dates", "matplotlib.colors", "matplotlib.cm",
    "netCDF4", "skimage", "skimage.transform",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "chips"))


def _install_stubs():
    for m in _STUBS:
        if m not in sys.modules:
            sys.modules[m] = MagicMock()


def load_reference():
    """Return (Chips, SynopticChips) classes of the reference, unmodified."""
    if not available():
        raise RuntimeError("reference tree not present")
    _install_stubs()
    try:
        from loguru import logger
        logger.disable("chips")
        logger.disable("syn_chips")
    except Exception:
        pass
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from chips.chips import Chips  # noqa
    p = os.path.join(REF, "chips")
    if p not in sys.path:
        sys.path.insert(0, p)          # bare `from plots import` at syn_chips.py:20
    from syn_chips import SynopticChips  # noqa
    return Chips, SynopticChips


def run_disk(disk, aia, base_folder="/tmp/chips_ref/", **kw):
    """Stage methods in `process_CHIPS` order (chips.py:152-171) minus the plotter."""
    Chips, _ = load_reference()
    c = Chips(aia, base_folder=base_folder, **kw)
    c.extract_solar_masks(disk)
    c.run_filters(disk)
    c.extract_histogram(disk)
    c.extract_CHs_CHBs(disk)
    return c


def run_synoptic(syn, base_folder="/tmp/chips_ref/", **kw):
    """Stage methods in `SynopticChips.run_CHIPS` order (syn_chips.py:79-93) minus plots."""
    _, SynopticChips = load_reference()
    c = SynopticChips(syn, base_folder=base_folder, **kw)
    c.run_filters(syn)
    c.extract_histogram(syn)
    c.extract_CHs_CHBs(syn)
    return c
