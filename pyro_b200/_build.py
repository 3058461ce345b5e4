"""Build the native library in-tree with nvcc for sm_100a.

``libpyro_b200.so`` (the C-ABI of include/pyro_b200.h; pure CUDA runtime, no torch types) and
``libpyro_b200_hostcheck.so`` (test-only CPU build of the element functors).  The objects go to
``build/`` and the shared objects next to this file, so they travel to the GPU box with the
snapshot (they are git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(ROOT, "build")
LIB = os.path.join(HERE, "libpyro_b200.so")
HOSTCHECK = os.path.join(HERE, "libpyro_b200_hostcheck.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                     "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets"]
# everything but the test harness
DEVICE_SOURCES_EXCLUDE = {"hostcheck.cu"}


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libpyro_b200.so")
    return exe


def _newest_header_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def _compile(src, obj, extra=()):
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    hdr_m = _newest_header_mtime()
    sources = sorted(f for f in os.listdir(CSRC)
                     if f.endswith(".cu") and f not in DEVICE_SOURCES_EXCLUDE)
    jobs = []
    objs = []
    for f in sources:
        src = os.path.join(CSRC, f)
        obj = os.path.join(BUILD, f[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or \
            os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m)
        if stale:
            jobs.append((src, obj))
    if jobs:
        if verbose:
            print("[pyro_b200] compiling %d CUDA sources for sm_100a" % len(jobs), file=sys.stderr)
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda so: _compile(*so), jobs))
    need_link = bool(jobs) or not os.path.exists(LIB) or \
        any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if need_link:
        cmd = [_nvcc()] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs + \
              ["-lcudart", "-Wno-deprecated-gpu-targets"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    # host-only harness (tests)
    hsrc = os.path.join(CSRC, "hostcheck.cu")
    if force or not os.path.exists(HOSTCHECK) or \
            os.path.getmtime(HOSTCHECK) < max(os.path.getmtime(hsrc), hdr_m):
        cmd = [_nvcc(), "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               "-Wno-deprecated-gpu-targets", hsrc, "-o", HOSTCHECK]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("hostcheck build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
