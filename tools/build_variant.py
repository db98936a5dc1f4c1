#!/usr/bin/env python3
"""Developer tool: build the library with extra -D flags into build/lib_<name>.so (select it with NRRT_B200_LIB).
Usage: build_variant.py <name> [-DMACRO=VALUE ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nr_ray_tracer_b200 import build as B  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build", f"lib_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = [B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ccbin", "/usr/bin/g++",
       "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-cudart", "static", "-shared", "-o", out] + flags + \
      [os.path.join(B.CSRC, f) for f in B.SOURCES]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode != 0:
    sys.exit("nvcc failed:\n" + r.stdout + r.stderr)
print(out)
