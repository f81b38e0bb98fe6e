"""Build libv2a_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("V2A_LIB") or os.path.join(PKG_DIR, "libv2a_b200.so")   # V2A_LIB: developer experiments
STAMP = os.path.join(PKG_DIR, ".libv2a_b200.stamp")

SOURCES = ["igemm.cu", "wgrad.cu", "elementwise.cu", "attention.cu", "policy.cu", "encoder.cu", "replay.cu", "perceiver.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--shared",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/v2a_b200.h"]
    for name in files:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into one shared library; returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *srcs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libv2a_b200.so:\n" + proc.stderr[-4000:])
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    with open(os.path.join(PKG_DIR, ".ptxas.log"), "w") as f:
        f.write(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_extension(force="--force" in sys.argv, verbose=True))
