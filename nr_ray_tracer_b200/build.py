"""Builds libnrrt_b200.so (CUDA kernels + C ABI + host scene layer) in-tree for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnrrt_b200.so")
CLI = os.path.join(HERE, "bin", "nr-ray-tracer")
SOURCES = ["nrrt_device.cu", "host_scene.cpp", "scene_loader.cpp"]
HEADERS = ["rt_device.cuh", "pool_kernel.cuh", "host_math.hpp", "jpeg_baseline.hpp", os.path.join("..", "..", "include", "nrrt.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    if not os.path.exists(CLI) or os.path.getmtime(CLI) < os.path.getmtime(os.path.join(CSRC, "cli_main.cpp")):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-ccbin", host_cxx, "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-cudart", "static",
           "-shared", "-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    build_cli(host_cxx)
    return LIB


def build_cli(host_cxx: str = "/usr/bin/g++") -> str:
    """`nr-ray-tracer render ...` (csrc/cli_main.cpp) linked against the in-tree library."""
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cmd = [host_cxx, "-O2", "-std=c++17", os.path.join(CSRC, "cli_main.cpp"), "-o", CLI, "-L" + HERE, "-lnrrt_b200",
           "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the CLI failed:\n" + r.stdout + r.stderr)
    return CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
