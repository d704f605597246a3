"""In-tree build of libdmk.so (hand-written sm_100a kernels + C ABI) with nvcc.

    python -m deepmimo_b200.build            # build if sources are newer than the library
    python -m deepmimo_b200.build --force

The library is built next to this file (deepmimo_b200/libdmk.so) so that it travels with the
repo snapshot to the GPU box; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libdmk.so")
SOURCES = ["dmk_api.cu"]
HEADERS = ["dmk_common.cuh", "dmk_prologue.cuh", "dmk_fd.cuh", "dmk_fd_tc.cuh", "dmk_fd_ws.cuh", "dmk_fd_small.cuh", "dmk_fd_mma.cuh", "dmk_fd_rows.cuh", "dmk_td.cuh", "dmk_bf.cuh", os.path.join("..", "..", "include", "dmk.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",                      # accumulation loops want FMA; the prologue uses explicit _rn intrinsics
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdmk.so cannot be built (no CPU fallback exists)")


STAMP = LIB + ".srchash"      # content hash of the sources the library was built from (mtimes do not survive a snapshot copy)


def source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for p in [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]:
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode() + b"\0" + f.read())
    h.update(os.environ.get("DMK_NVCC_EXTRA", "").encode())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile deepmimo_b200/csrc/*.cu into deepmimo_b200/libdmk.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    # One builder at a time (every torchrun rank may find the library stale at once): an exclusive file lock, a per-process
    # temporary output, and a re-check under the lock so that the ranks that waited reuse the winner's build.
    import fcntl
    with open(os.path.join(PKG, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():
            return LIB
        extra = os.environ.get("DMK_NVCC_EXTRA", "").split()      # e.g. -DDMK_TC_TRACE for the phase-timing debug build
        tmp = f"{LIB}.{os.getpid()}.tmp"
        cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", tmp, *[os.path.join(CSRC, s) for s in SOURCES]]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log = proc.stdout + proc.stderr
        with open(os.path.join(PKG, "build.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if proc.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed:\n" + log)
        os.replace(tmp, LIB)
        with open(STAMP, "w") as f:
            f.write(source_hash())
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
