"""In-tree build of libsupernet_b200.so (the C-ABI of include/supernet.h) with nvcc for sm_100a.

The library links only the CUDA runtime (static) and resolves the driver's tensor-map encoders at run
time through cudaGetDriverEntryPoint, so it loads (and its symbols can be listed) on a box without a
GPU driver.  It does NOT link libtorch: PyTorch reaches it through ctypes with raw pointers.
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_NAME = "libsupernet_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
STAMP = os.path.join(PKG_DIR, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
]
if os.environ.get("SN_BUILD_KNOBS") == "1":          # profiling builds only: compile the SN_HL_DBG knob checks in
    NVCC_FLAGS.append("-DSN_HL_KNOBS")


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build libsupernet_b200.so")
    return cand


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h"))):
        # file NAME, not path: the stamp must stay valid when the tree is copied elsewhere (the GPU box runs from a
        # scratch copy; a path-dependent digest made every process there rebuild, and eight ranks doing so at once
        # corrupted each other's object files)
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link the shared library next to this file."""
    if not force and not needs_build():
        return LIB_PATH
    import fcntl
    lock_path = os.path.join(PKG_DIR, ".build_lock")
    with open(lock_path, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)          # one builder at a time (torchrun starts one process per GPU)
        try:
            if not force and not needs_build():   # another process built it while we waited
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out, file=sys.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    tmp_lib = LIB_PATH + f".tmp{os.getpid()}"
    link = [nvcc, "-shared", "-o", tmp_lib, *objs, "-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp_lib, LIB_PATH)                 # atomic: a concurrent loader never sees a half-written library
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
