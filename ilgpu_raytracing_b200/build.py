"""Builds the native libraries in-tree with nvcc for sm_100a (no JIT cache, so the .so travels to the GPU box)."""
from __future__ import annotations

import glob
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              # exact-arithmetic contract of rt_core.h: no contraction, IEEE div/sqrt, denormals kept
              "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-shared"]

CORE_SO = os.path.join(_PKG, "librtcore_b200.so")
CORE_SRCS = [os.path.join(_CSRC, f) for f in ("rtcore.cu", "rt_bvh.cpp")]
CORE_DEPS = CORE_SRCS + sorted(glob.glob(os.path.join(_CSRC, "*.h"))) + [os.path.join(_PKG, "..", "include", "rtcore_b200.h")]   # every header counts

ENGINE_SO = os.path.join(_PKG, "librtengine_host.so")
ENGINE_SRCS = [os.path.join(_CSRC, "host", f) for f in ("engine.cpp", "mesh_loader_obj.cpp", "png_decode.cpp")]
ENGINE_DEPS = ENGINE_SRCS + [os.path.join(_CSRC, "host", "engine.h"), os.path.join(_PKG, "..", "include", "rtcore_b200.h")]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_core(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(CORE_SO, CORE_DEPS):
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", CORE_SO] + CORE_SRCS
        subprocess.run(cmd, check=True)
    return CORE_SO


def build_engine(force: bool = False) -> str:
    if not all(os.path.exists(s) for s in ENGINE_SRCS):
        return ""
    if force or _stale(ENGINE_SO, ENGINE_DEPS):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-pthread",
               "-o", ENGINE_SO] + ENGINE_SRCS + ["-L" + _PKG, "-l:librtcore_b200.so", "-Wl,-rpath,$ORIGIN"]
        subprocess.run(cmd, check=True)
    return ENGINE_SO


def build_all(force: bool = False) -> None:
    build_core(force)
    build_engine(force)


if __name__ == "__main__":
    build_all(force=True)
