"""In-tree build of libtai_b200.so (sm_100a only).

``python -m video_frame_inpainting_b200.build`` or ``build_library()``: every ``csrc/*.cu`` is compiled
with ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` (nvcc cross-compiles without a GPU)
and linked into ``video_frame_inpainting_b200/lib/libtai_b200.so``.  The .so is git-ignored but
travels to the GPU box with the repository snapshot.

Replaces the reference's two-step build (bashes/misc/install.bash:3-8 -> nvcc -c with a user-given
arch, then src/separable_convolution/install.py:11-32 -> torch.utils.ffi.create_extension).
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtai_b200.so")
HASH_PATH = LIB_PATH + ".srchash"
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    # no --use_fast_math: the warp / blend kernels rely on IEEE rounding of every operation (floor() of the
    # coordinate chain is bit-exact against the FP32 reference) and the separable convolutions on plain FMA;
    # the gate kernels pick their own approximations explicitly (ex2.approx / rcp.approx, elementwise.cu)
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found: libtai_b200.so cannot be built")
    return exe


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h")))


def source_hash() -> str:
    """Content hash of every source the library is built from (mtimes do not survive the copy to
    the GPU box, contents do)."""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in _deps():
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_up_to_date() -> bool:
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(HASH_PATH)):
        return False
    with open(HASH_PATH) as f:
        return f.read().strip() == source_hash()


def build_library(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """Compile and link the library if any source is newer than it.  Returns the .so path."""
    if not force and is_up_to_date():
        return LIB_PATH
    nvcc = _nvcc()
    started_from = source_hash()   # recorded at the end: an edit DURING the build leaves the library marked stale
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    hdr_mtime = max([os.path.getmtime(p) for p in glob.glob(os.path.join(CSRC, "*.cuh")) +
                     glob.glob(os.path.join(INCLUDE, "*.h"))] or [0.0])

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.isfile(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_mtime)):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
        if verbose and res.stderr.strip():
            print(res.stderr, flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, _sources()))
    tmp = LIB_PATH + ".tmp"
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", tmp],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    os.replace(tmp, LIB_PATH)
    with open(HASH_PATH, "w") as f:
        f.write(started_from + "\n")
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose=True)
    print(path)
