"""In-tree build of libt4b.so (sm_100a only) and of the test/oracle helper libraries.

    python tensor4all-rs_b200/build.py            # product library
    python tensor4all-rs_b200/build.py --oracle   # + oracle/_build/liboracle.so (C restatement)

nvcc cross-compiles without a GPU.  Objects are cached under build/ by mtime.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
              "--expt-relaxed-constexpr"] + ARCH
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-Wall"]


def _sources():
    cu, cpp = [], []
    for base, _, files in os.walk(CSRC):
        for f in sorted(files):
            p = os.path.join(base, f)
            if f.endswith(".cu"):
                cu.append(p)
            elif f.endswith(".cpp"):
                cpp.append(p)
    return cu, cpp


def _headers_mtime():
    m = 0.0
    for base, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(base, f)))
    m = max(m, os.path.getmtime(os.path.join(ROOT, "include", "t4b.h")))
    return m


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + os.path.basename(cmd[-1]))
    return r.stdout + r.stderr


def _obj_for(src, tag=""):
    rel = os.path.relpath(src, ROOT).replace(os.sep, "_")
    return os.path.join(BUILD, tag + rel + ".o")


def _compile(src, obj, extra, hmtime, verbose):
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hmtime):
        return ""
    if src.endswith(".cu"):
        cmd = [NVCC] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
    else:
        cmd = [NVCC, "-std=c++17", "-O2", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"] + extra + [
            "-c", src, "-o", obj]
    out = _run(cmd)
    if verbose:
        print("compiled", os.path.relpath(src, ROOT))
    return out


def build_product(verbose=True):
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    cu, cpp = _sources()
    hm = _headers_mtime()
    objs = []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        futs = []
        for s in cu + cpp:
            o = _obj_for(s)
            objs.append(o)
            futs.append(ex.submit(_compile, s, o, [], hm, verbose))
        for f in futs:
            f.result()
    lib = os.path.join(LIBDIR, "libt4b.so")
    if (not os.path.exists(lib)) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
        # static cudart: the library carries its own runtime and shares the primary context
        # with torch, so device pointers / streams from torch are directly usable.
        _run([NVCC] + ARCH + ["-shared", "-cudart", "static", "-o", lib] + objs + ["-lpthread"])
        if verbose:
            print("linked", os.path.relpath(lib, ROOT))
    return lib


def build_oracle(verbose=True):
    odir = os.path.join(ROOT, "oracle")
    out_dir = os.path.join(odir, "_build")
    os.makedirs(out_dir, exist_ok=True)
    srcs = [os.path.join(odir, f) for f in sorted(os.listdir(odir)) if f.endswith(".c")]
    lib = os.path.join(out_dir, "liboracle.so")
    if not srcs:
        return None
    if (not os.path.exists(lib)) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        # -ffp-contract=off: the rrLU restatement must not fuse `t - x*y` (bit-exact pivots)
        _run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-o", lib] + srcs + ["-lm"])
        if verbose:
            print("linked", os.path.relpath(lib, ROOT))
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--all", action="store_true")
    a = ap.parse_args()
    build_product()
    if a.oracle or a.all:
        build_oracle()
