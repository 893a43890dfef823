"""Builds liblhgt.so (sm_100a kernels + C ABI) and the `extract_ref` executable in-tree with nvcc.

The artefacts land in localhgt_b200/_build/ (git-ignored, but they travel to the GPU box with the
gpurun snapshot).  No JIT, no torch extension machinery: the product is a plain C-ABI shared library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "liblhgt.so")
EXE = os.path.join(OUT, "extract_ref")
EXE_REGIONS = os.path.join(OUT, "extract_regions")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("lhgt_kernels.cu", "lhgt_api.cu")]
    deps = srcs + [os.path.join(CSRC, "lhgt_kernels.cuh"), os.path.join(HERE, "..", "include", "lhgt.h"),
                   os.path.abspath(__file__)]
    if force or _stale(LIB, deps):
        cmd = [_nvcc(), "-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-Wall", "-shared",
               "-Xptxas", "-v" if verbose else "-warn-spills", "-o", LIB, *srcs]
        subprocess.check_call(cmd)
    main_src = os.path.join(CSRC, "extract_ref_main.cpp")
    if force or _stale(EXE, [main_src, LIB]):
        subprocess.check_call(["g++", "-O2", "-o", EXE, main_src, "-L" + OUT, "-llhgt", "-Wl,-rpath,$ORIGIN"])
    regions_src = os.path.join(CSRC, "extract_regions_main.cpp")
    if force or _stale(EXE_REGIONS, [regions_src, LIB]):
        subprocess.check_call(["g++", "-O2", "-o", EXE_REGIONS, regions_src, "-L" + OUT, "-llhgt", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
