"""In-tree build of libcd360.so (sm_100a only) with plain nvcc.

The library is the C-ABI boundary declared in include/cd360.h; it has no torch / Python
dependency.  `build()` is what `__graft_entry__.build()` calls; it cross-compiles on a CPU-only
box (nvcc does not need a GPU) and leaves the .so next to this file so it travels with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libcd360.so")

SOURCES = [
    "gemm_tcgen05.cu",
    "attention_tcgen05.cu",
    "norm.cu",
    "elementwise.cu",
    "nerf.cu",
    "attention_bwd.cu",
    "attention_bwd_tcgen05.cu",
    "train.cu",
    "vae.cu",
    "conditioner.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libcd360.so")


HASH_FILE = LIB + ".srchash"


def _src_hash() -> str:
    """Content hash of every input of the build (mtimes do not survive the copy to the GPU box)."""
    import hashlib

    h = hashlib.sha256()
    paths = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    paths.append(os.path.join(HERE, "..", "include", "cd360.h"))
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(HASH_FILE):
        return True
    with open(HASH_FILE) as f:
        return f.read().strip() != _src_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libcd360.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    hdr_mtime = max(
        os.path.getmtime(os.path.join(CSRC, "cd360_common.cuh")),
        os.path.getmtime(os.path.join(HERE, "..", "include", "cd360.h")),
    )

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        srcp = os.path.join(CSRC, src)
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), hdr_mtime)):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", srcp, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(BUILD, src + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB + ".tmp", *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(LIB + ".tmp", LIB)
    with open(HASH_FILE, "w") as f:
        f.write(_src_hash())
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose=True))
