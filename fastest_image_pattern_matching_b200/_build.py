"""Builds libfpm_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfpm_b200.so")
SOURCES = ["fpm_host.cu"]
HEADERS = ["fpm_common.cuh", "fpm_geometry.cuh", "fpm_kernels.cuh", "fpm_pyrdown.cuh", "fpm_mma.cuh", "fpm_fused.cuh", "fpm_jpeg.h"]
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # epilogues must round like the reference's non-FMA x86-64 build
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(HERE, "..", "include", "fpm_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library if it is missing or older than its sources; returns its path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def build_pybind(force: bool = False) -> str:
    """Compile the pybind11 module fpm_b200_pybind (g++, links libfpm_b200.so by rpath $ORIGIN)."""
    import sysconfig
    import pybind11
    build()
    out = os.path.join(HERE, "fpm_b200_pybind" + sysconfig.get_config_var("EXT_SUFFIX"))
    src = os.path.join(CSRC, "fpm_pybind.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "fpm_template_matcher.hpp"), os.path.join(HERE, "..", "include", "fpm_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
           src, "-o", out, "-L", HERE, "-l:libfpm_b200.so", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ (pybind) failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
