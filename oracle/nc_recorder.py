"""TEST INFRASTRUCTURE — not product code.

A recording stand-in for `netCDF4.Dataset` (the package is absent from this image): every
createGroup / createDimension / createVariable / assignment is logged, with a digest of the values
cast to the variable's dtype.  `tests/golden/make_golden.py` runs the UNMODIFIED reference's
`Chips.to_netcdf` (/root/reference/chips/chips.py:512-590) against it and commits the log
(`tests/golden/netcdf_layout_n256.json`); the GPU test runs `pychips_b200.Chips.to_netcdf`'s NETCDF4
branch against the same recorder and compares the two logs line by line.
"""
from __future__ import annotations

import hashlib
import os
import types

import numpy as np

LAST = []          # the datasets opened so far (the newest last)


def _dims(d):
    return [d] if isinstance(d, str) else list(d)


class _Var:
    def __init__(self, ds, group, name, dtype):
        self.ds, self.group, self.name, self.dtype = ds, group, name, dtype

    def __setitem__(self, key, value):
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float64).astype(np.dtype(self.dtype)))
        self.ds.log.append(["data", self.group, self.name, list(a.shape), str(a.dtype),
                            hashlib.sha256(a.tobytes()).hexdigest()])


class _Group:
    def __init__(self, ds, name):
        self.ds, self.name = ds, name

    def createDimension(self, name, size):
        self.ds.log.append(["dim", self.name, name, int(size)])

    def createVariable(self, name, dtype, dims):
        self.ds.log.append(["var", self.name, name, dtype, _dims(dims)])
        return _Var(self.ds, self.name, name, dtype)


class Dataset:
    def __init__(self, path, mode="r", format=None):
        self.log = [["open", os.path.basename(path), mode, format]]
        LAST.append(self)

    def createGroup(self, name):
        self.log.append(["group", name])
        return _Group(self, name)

    def close(self):
        self.log.append(["close"])


def as_module():
    m = types.ModuleType("netCDF4")
    m.Dataset = Dataset
    return m
