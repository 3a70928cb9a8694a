"""Build libmethylasso_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile).

    python -m methylasso_b200.build [--force]

The library is the product: hand-written CUDA kernels + C++ host orchestration behind the C ABI of
include/methylasso_b200.h.  Flags that matter for parity: -fmad=false (the reference is built
without FMA contraction) and no fast-math (zero weights rely on IEEE inf).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("ML_BUILD_OUT") or os.path.join(HERE, "libmethylasso_b200.so")   # (experiments build variants elsewhere)
SOURCES = ["flsa.cu", "detect.cu", "segment.cu", "stats.cu", "codec.cu", "capi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _newest_source():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(HERE, "..", "include", "methylasso_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.environ.get("ML_BUILD_OBJDIR") or os.path.join(HERE, "_build")
    extra = os.environ.get("ML_BUILD_DEFS", "").split()
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    # host-only passes over the input columns: the host compiler alone (function clones for AVX2, chosen at load time)
    hobj = os.path.join(objdir, "hostpack.o")
    hcmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-pthread", "-c",
            os.path.join(CSRC, "hostpack.cpp"), "-o", hobj]
    if verbose:
        print(" ".join(hcmd))
    subprocess.check_call(hcmd)
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out.decode())
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", OUT, *objs, hobj, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-pthread"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
