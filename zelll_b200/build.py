"""In-tree build of libzelll_b200.so (nvcc, sm_100a only).

`python -m zelll_b200.build` or `zelll_b200.build.build()`.  The .so is git-ignored but travels
to the GPU box with the snapshot.  Flags: -fmad=false / -prec-div=true keep every floating-point
operation a separately rounded IEEE operation, as in the Rust reference.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzelll_b200.so")
SOURCES = ["zelll_b200.cu"]
HEADERS = ["common.cuh", "build_kernels.cuh", "pair_kernels.cuh", "pair_pf_kernels.cuh", "p2p_kernels.cuh", "query_kernels.cuh", "sparse_kernels.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libzelll_b200.so")


def source_hash() -> str:
    """SHA-256 over the compiler flags and every source the library is built from (names and bytes, in
    a fixed order).  Compiled in as zb_build_id(); tests compare the two, so a stale or foreign .so is
    detected instead of being tested by accident."""
    import hashlib

    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "zelll_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def built_hash() -> str:
    """zb_build_id() of the library on disk ('' when it is missing or predates the symbol)."""
    if not os.path.exists(LIB):
        return ""
    with open(LIB, "rb") as f:
        blob = f.read()
    k = blob.find(b"zb-build-id:")
    return blob[k + 12:k + 12 + 64].decode("ascii", "replace") if k >= 0 else ""


def stale() -> bool:
    return built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, f'-DZB_BUILD_ID="zb-build-id:{source_hash()}"', "-o", LIB]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
