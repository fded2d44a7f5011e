"""Builds gnomix_b200/libgnx.so (hand-written CUDA for sm_100a) in-tree with nvcc.

Called by __graft_entry__.build(); nvcc cross-compiles without a GPU.  The built
library travels with the repo snapshot to the GPU box (it is git-ignored, not
gpurun-ignored).
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgnx.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "-cudart", "static",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _deps():
    inc = os.path.join(HERE, "..", "include")
    return (sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
            + glob.glob(os.path.join(inc, "*.h")))


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libgnx.so")
    objdir = os.path.join(HERE, "..", "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.splitext(os.path.basename(src))[0] + ".o")
        objs.append(obj)
        if (not force) and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(p) for p in _deps() if not p.endswith((".cu", ".cpp")) or p == src):
            continue
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", src, "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
        if verbose and out:
            print(out)
    tmp = LIB + ".tmp"
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs + ["-lpthread", "-lz"]
    subprocess.check_call(cmd, env=env)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
