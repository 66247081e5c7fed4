"""Builds aero_b200/libaero_b200.so (CUDA kernels + C ABI + host driver) for sm_100a, in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["csrc/abi.cu", "csrc/ntt.cu", "csrc/hash.cu", "csrc/poly.cu", "csrc/fri.cu", "csrc/peak.cu", "host/prover.cpp"]
HEADERS = ["csrc/gl.cuh", "csrc/blake2s.cuh", "csrc/kernels.cuh", "csrc/ntt.cuh", "host/prover.hpp", "host/copy_pool.hpp",
           "../include/aero_b200.h", "../include/aero_prover.h"]
LIB = os.path.join(_HERE, "libaero_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-x", "cu"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the aero_b200 CUDA library cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(_HERE, f)) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc; let nvcc pick the system g++
    env.pop("CXX", None)
    subprocess.check_call(cmd, cwd=_HERE, env=env)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=False))
