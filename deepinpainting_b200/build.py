"""Build libipsr_sm100.so (sm_100a only) in-tree with nvcc.

    python -m deepinpainting_b200.build [--force] [--verbose]

The shared library lands in ``deepinpainting_b200/lib/`` (git-ignored, but it travels to the GPU
box with the repository snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "lib", "obj")
LIBNAME = "libipsr_sm100.so"
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libipsr_sm100.so cannot be built (no CPU fallback exists)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    files.append(os.path.join(INCLUDE, "ipsr_sm100.h"))
    # names relative to the package and flags without the absolute include path: the snapshot that travels to a GPU box
    # lives under another directory there, and must not look stale for that alone
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(a for a in NVCC_FLAGS if a != INCLUDE).encode())
    return h.hexdigest()


def library_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def is_current() -> bool:
    stamp = os.path.join(LIBDIR, "build.sha256")
    if not (os.path.exists(library_path()) and os.path.exists(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == _digest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and is_current():
        return library_path()
    # one builder at a time (the ranks of a torchrun / mp.spawn job may all find the library missing or stale at once); the
    # others wait and then find it current
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():
                return library_path()
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as pool:
        objs = list(pool.map(compile_one, sources()))
    tmp = library_path() + ".tmp.%d" % os.getpid()        # link beside the target, then swap in atomically
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    os.replace(tmp, library_path())
    with open(os.path.join(LIBDIR, "build.sha256"), "w") as fh:
        fh.write(_digest())
    return library_path()


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
