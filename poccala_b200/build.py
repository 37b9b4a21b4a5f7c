"""Build the sm_100a shared library behind include/poccala_b200.h (in-tree, nvcc only).

    python -m poccala_b200.build            # incremental
    python -m poccala_b200.build --force

Output: poccala_b200/_lib/libpoccala_b200.so (git-ignored; travels to the GPU box with the tree).
nvcc cross-compiles without a GPU.  No torch headers are involved: the ABI is plain C.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libpoccala_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
         "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the engine cannot be built")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "poccala_b200.h"))
    return max(os.path.getmtime(p) for p in hdrs)


def _compile(src, obj, verbose):
    cmd = [nvcc(), *ARCH, *FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, log))
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sources()
    hdr_m = _deps_mtime()
    objs, todo = [], []
    for s in srcs:
        o = os.path.join(LIBDIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            todo.append((s, o))
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda so: _compile(so[0], so[1], verbose), todo))
    if todo or not os.path.exists(LIB):
        cmd = [nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
