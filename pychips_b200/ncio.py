"""Minimal NetCDF-3 (64-bit offset, "CDF\\x02") writer for fixed-size float32 variables.

`Chips.to_netcdf` (`/root/reference/chips/chips.py:512-590`) uses the netCDF4 package; when it
is not installed the same arrays are written in the classic format, which needs no third-party
code: a header (dimension list, variable list) followed by the big-endian variable data.
Classic files have no groups, so callers name variables "<group>_<variable>".
"""
from __future__ import annotations

import struct

import numpy as np

_NC_DIMENSION, _NC_VARIABLE, _NC_FLOAT = 0x0A, 0x0B, 5


def _name(s: str) -> bytes:
    b = s.encode("ascii")
    return struct.pack(">i", len(b)) + b + b"\0" * (-len(b) % 4)


def write_float32(path: str, dims: dict, variables: dict) -> None:
    """dims: name -> size (> 0); variables: name -> (tuple of dim names, array)."""
    dim_names = list(dims)
    head = b"CDF\x02" + struct.pack(">i", 0)
    head += struct.pack(">ii", _NC_DIMENSION, len(dims)) if dims else struct.pack(">ii", 0, 0)
    for d in dim_names:
        head += _name(d) + struct.pack(">i", int(dims[d]))
    head += struct.pack(">ii", 0, 0)                          # no global attributes
    entries = []
    for vname, (vdims, value) in variables.items():
        shape = tuple(int(dims[d]) for d in vdims)
        arr = np.asarray(value, dtype=np.float32)
        if arr.shape != shape:
            raise ValueError(f"{vname}: shape {arr.shape} does not match dimensions {shape}")
        nbytes = int(np.prod(shape)) * 4
        entries.append((vname, vdims, arr, nbytes))
    fixed = struct.pack(">ii", _NC_VARIABLE, len(entries)) if entries else struct.pack(">ii", 0, 0)
    # variable records have a fixed size once the names are known: compute the data offsets
    recs = []
    for vname, vdims, _arr, _nb in entries:
        recs.append(_name(vname) + struct.pack(">i", len(vdims)) +
                    b"".join(struct.pack(">i", dim_names.index(d)) for d in vdims) +
                    struct.pack(">ii", 0, 0) + struct.pack(">i", _NC_FLOAT))
    offset = len(head) + len(fixed) + sum(len(r) + 4 + 8 for r in recs)
    body = b""
    for rec, (_vn, _vd, _arr, nbytes) in zip(recs, entries):
        vsize = nbytes if nbytes < 2 ** 32 - 4 else 2 ** 32 - 1
        body += rec + struct.pack(">I", vsize) + struct.pack(">q", offset)
        offset += nbytes
    with open(path, "wb") as f:
        f.write(head + fixed + body)
        for _vn, _vd, arr, _nb in entries:
            arr.astype(">f4", copy=False).tofile(f) if arr.dtype.byteorder == ">" else arr.astype(">f4").tofile(f)
