"""Builds libcrb3d_sm100.so (the C-ABI library of include/crb3d.h) in-tree with nvcc for sm_100a only.

Usage: python crb-active-3ddet_b200/build.py [--force]
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcrb3d_sm100.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-diag-suppress", "177"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, ".digest")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    # headers are shared: a source is recompiled when it, any header or the flags changed (per-object stamp files)
    hh = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")):
            hh.update(open(os.path.join(CSRC, f), "rb").read())
    hh.update(" ".join(ARCH + FLAGS).encode())
    hdr_dig = hh.hexdigest()

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        d = hashlib.sha256(open(os.path.join(CSRC, src), "rb").read() + hdr_dig.encode()).hexdigest()
        st = obj + ".digest"
        if not force and os.path.exists(obj) and os.path.exists(st) and open(st).read() == d:
            return obj
        cmd = [NVCC, *ARCH, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        open(st, "w").write(d)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(dig)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
