"""Build the C-ABI CUDA library (libmixdq_b200.so) in-tree with nvcc for sm_100a.

The library has no torch / Python dependency: plain `extern "C"` entry points declared in
include/mixdq_b200.h. It is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libmixdq_b200.so"
STAMP = PKG_DIR / ".libmixdq_b200.stamp"
SOURCES = ["quant.cu", "quant2.cu", "fused_quant.cu", "simt.cu", "capi.cu", "persist.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
    "--threads", "8",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libmixdq_b200.so")


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*")) + [PKG_DIR.parent / "include" / "mixdq_b200.h"]
    for f in files:
        if f.is_file():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if sources changed. Returns the path of the .so."""
    want = _source_hash()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == want:
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(CSRC / s) for s in SOURCES] + ["-o", str(LIB_PATH)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libmixdq_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    STAMP.write_text(want)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
