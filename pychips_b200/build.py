"""Build the C-ABI shared library `libchips_b200.so` in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("CHIPS_LIB", os.path.join(HERE, "libchips_b200.so"))   # override: diagnostics only
SOURCES = ["median.cu", "stats.cu", "ccl.cu", "fl.cu", "export.cu", "contours.cu", "pipeline.cu", "batch.cu", "synccl.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math=false", "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "chips_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("CHIPS_EXTRA_NVCC_FLAGS", "").split()
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", os.environ.get("CHIPS_LIB_OUT", LIB)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
