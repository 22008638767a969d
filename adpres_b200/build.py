"""Build recipe for libadpres_b200.so (CUDA kernels + C ABI + C++ host mirror), in-tree.

sm_100a only.  -fmad=false keeps every product/sum rounded like the reference build
(gfortran -O4, x86-64, no FMA contraction): the kernels are HBM-bound, so fusing
multiply-adds would buy nothing and would move results off the reference by 1 ulp per
operation.  -lineinfo keeps the ncu source page usable.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libadpres_b200.so")
SOURCES = ["cmfd_kernels.cu", "nodal_kernels.cu", "capi.cu", "results.cu", "th.cu", "comm.cu", "host_cmfd.cpp"]
import glob
# every header of csrc/ is a dependency of every object (edits to any .cuh rebuild the library)
HEADERS = sorted(os.path.basename(h) for h in glob.glob(os.path.join(CSRC, "*.cuh"))) + \
          [os.path.join("..", "..", "include", "adpres_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-ffp-contract=off", "-Xptxas", "-v"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc()] + NVCC_FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", s, "-o", o]
            procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, cmd, pr in procs:
        out, _ = pr.communicate()
        with open(os.path.join(objdir, os.path.splitext(src)[0] + ".log"), "w") as fh:
            fh.write(" ".join(cmd) + "\n" + out)
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            print(out)
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-ldl", "-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
