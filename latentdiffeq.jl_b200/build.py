"""Builds libldeq.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python latentdiffeq.jl_b200/build.py [--force] [--verbose]

The library has no torch dependency: it links the CUDA runtime statically and dlopen()s NVRTC
lazily for user-defined right-hand sides.  The built .so is git-ignored but travels with gpurun.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libldeq.so")
SOURCES = ["ldeq_api.cu", "ldeq_loss.cu", "ldeq_mlp.cu", "ldeq_user_rhs.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _host_cxx() -> list[str]:
    # the image's $CXX wrapper misses pieces; /usr/bin/g++ is complete
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ldeq.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None, out: str | None = None) -> str:
    """``extra``/``out`` build a tuning variant (other -D flags) next to the product library."""
    LIB = out or globals()["LIB"]
    if not force and not out and not _stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objdir = os.path.join(HERE, "build", os.path.basename(LIB))
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    common = [nvcc, *ARCH, *_host_cxx(), "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-I", os.path.join(HERE, "..", "include"), "-I", CSRC, *(extra or [])]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen(common + ["-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} (exit {p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed; see log above")
    # static cudart (nvcc default); NVRTC is dlopen'ed lazily by the user-RHS path, so the library
    # loads on a box without a driver or GPU (the CPU test tier checks its exported symbols there)
    link = [nvcc, *ARCH, *_host_cxx(), "-shared", "-o", LIB, *objs, "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
