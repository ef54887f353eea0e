"""Builds the engine's shared library in-tree with nvcc for sm_100a.

    python -m snickery_b200.build [--force]

The .so lands in snickery_b200/_lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libsnk_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ["api.cu", "weights.cu", "knn_simt.cu", "knn_tc.cu", "rerank.cu", "search.cu", "join_viterbi.cu", "scores.cu", "concat.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "snk_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + \
          ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
