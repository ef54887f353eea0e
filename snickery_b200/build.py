"""Builds the engine's shared library in-tree with nvcc for sm_100a.

    python -m snickery_b200.build [--force]

Every .cu is compiled to an object of its own (in parallel, only when it or a header changed) and the objects are
linked into snickery_b200/_lib/libsnk_b200.so (git-ignored, shipped to the GPU box by gpurun).  NCCL is NOT a
link-time dependency: comm.cu resolves libnccl.so.2 with dlopen when snk_comm_init is first called.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libsnk_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ["api.cu", "weights.cu", "knn_simt.cu", "knn_tc.cu", "rerank.cu", "search.cu", "join_viterbi.cu", "scores.cu",
           "concat.cu", "comm.cu", "join_tc.cu", "greedy_one.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _obj(src):
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def _obj_stale(src, force):
    o = _obj(src)
    if force or not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(d) > t for d in [os.path.join(CSRC, src)] + _headers())


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources()] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [s for s in _sources() if _obj_stale(s, force)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", _obj(src)]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as pool:
        list(pool.map(compile_one, todo))
    tmp = LIB + ".tmp%d" % os.getpid()      # link beside the target, then rename: a reader never sees a half-written .so
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + [_obj(s) for s in _sources()] + \
          ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
