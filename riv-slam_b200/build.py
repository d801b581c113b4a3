"""Builds the product library ``libapdgicp_b200.so`` (hand-written sm_100a kernels + C ABI) in-tree.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting ``.so``
travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libapdgicp_b200.so")
HASH_PATH = LIB_PATH + ".srchash"
SOURCES = ["apd_build.cu", "apd_knn_cov.cu", "apd_align.cu", "apd_preprocess.cu", "apd_capi.cu"]
HEADERS = ["apd_internal.h", "apd_grid.cuh", "apd_leaf.cuh", "apd_merge_net.cuh", "apd_math.cuh", os.path.join("..", "..", "include", "apdgicp_b200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "--threads", "4",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the product has no CPU path and cannot be built without the CUDA toolkit")


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in SOURCES + HEADERS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    """Stale = the library is missing or was built from other sources / flags. The decision uses a content hash
    stored beside the library, not mtimes: a snapshot copied to another machine does not keep a usable mtime order."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != _source_hash()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Build the library if it is missing or older than its sources. Safe to call from several processes at once
    (one rank per GPU): the build runs under an exclusive file lock into a temporary file that is renamed into
    place, so a concurrent dlopen never sees a half-written library and only one process runs nvcc."""
    if not force and not needs_build():
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while this one waited
                return LIB_PATH
            tmp = f"{LIB_PATH}.tmp.{os.getpid()}"
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            # the image's default host compiler wrapper (/opt/gcc) lacks libgomp specs; the distro g++ works everywhere
            if os.path.exists("/usr/bin/g++"):
                cmd += ["-ccbin", "/usr/bin/g++"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
            os.replace(tmp, LIB_PATH)
            with open(HASH_PATH + ".tmp", "w") as f:
                f.write(_source_hash())
            os.replace(HASH_PATH + ".tmp", HASH_PATH)
            if verbose:
                print(r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
