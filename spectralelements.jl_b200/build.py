"""Build libsemb.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python spectralelements.jl_b200/build.py [--force] [--jobs N]

Objects go to spectralelements.jl_b200/_build/, the library to spectralelements.jl_b200/lib/libsemb.so
(git-ignored, but shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsemb.so")
STRIP_N = list(range(2, 18))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"] + \
    os.environ.get("SEMB_EXTRA_FLAGS", "").split()


def _sources_hash() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".cu", ".cuh", ".cpp", ".h")):
                with open(os.path.join(root, fn), "rb") as f:
                    h.update(fn.encode())
                    h.update(f.read())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def _run(cmd):
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), p.stdout, p.stderr))
    return p.stdout + p.stderr


def build(force: bool = False, jobs: int | None = None, verbose: bool = False, tag: str = "") -> str:
    """tag: build an experimental variant (SEMB_EXTRA_FLAGS=...) beside the product library, into lib/libsemb_<tag>.so
    with its own object directory; load it with SEMB_LIB=<path> (tools/ A/B scripts)."""
    global BUILD, LIB
    if tag:
        BUILD = os.path.join(HERE, "_build_" + tag)
        LIB = os.path.join(LIBDIR, "libsemb_%s.so" % tag)
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    want = _sources_hash()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB
    tasks = []
    objs = []
    for name in ("semb_api.cu", "semb_vec.cu", "semb_advect_tile.cu", "semb_stokes.cu", "semb_stokes_tile.cu", "semb_fdm.cu"):
        o = os.path.join(BUILD, name.replace(".cu", ".o"))
        objs.append(o)
        tasks.append([NVCC] + ARCH + COMMON + ["-c", os.path.join(CSRC, name), "-o", o])
    o = os.path.join(BUILD, "semb_host.o")
    objs.append(o)
    tasks.append([NVCC] + COMMON + ["-c", os.path.join(CSRC, "semb_host.cpp"), "-o", o])
    for n in STRIP_N:
        o = os.path.join(BUILD, "semb_strip_n%d.o" % n)
        objs.append(o)
        tasks.append([NVCC] + ARCH + COMMON + ["-DSEMB_INST_N=%d" % n, "-c",
                                              os.path.join(CSRC, "semb_strip_inst.cu"), "-o", o])
    jobs = jobs or min(len(tasks), os.cpu_count() or 4)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        for out in ex.map(_run, tasks):
            if verbose and out.strip():
                print(out)
    _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lnccl"])
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--tag", default="", help="experimental variant: lib/libsemb_<tag>.so (use with SEMB_EXTRA_FLAGS)")
    a = ap.parse_args()
    print(build(a.force, a.jobs, a.verbose, a.tag))
