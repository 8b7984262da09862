"""Builds libeconpdes_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python econpdes.jl_b200/build.py [--force] [--verbose]

The model/stencil/Jacobian translation units are compiled with -fmad=false: Julia never
contracts a*b+c into an fma, and residual parity with the CPU path is checked to the ulp.
The linear-solver units (krylov.cu) may contract.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libeconpdes_cuda.so")
OBJDIR = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
          "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
          "-DECONPDES_BUILD"]
MODELOPS_PARTS = 8   # kModelOpsParts (csrc/engine.cuh): slices of the functor registry, compiled in parallel
# object name -> (source, extra flags)
UNITS = {
    "engine.cu": ("engine.cu", ["-fmad=false"]),
    "api.cu": ("api.cu", ["-fmad=false"]),
    "krylov.cu": ("krylov.cu", []),
    "dist.cu": ("dist.cu", []),
    "pattern.cpp": ("pattern.cpp", []),
    "errors.cpp": ("errors.cpp", []),
}
for _k in range(MODELOPS_PARTS):
    UNITS["modelops%d.cu" % _k] = ("modelops.cu", ["-fmad=false", "-DEP_PART=%d" % _k])


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(os.path.dirname(HERE), "include", "econpdes_cuda.h")]


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(h) for h in _headers())
    objs, procs = [], []
    for unit, (source, extra) in UNITS.items():
        src = os.path.join(CSRC, source)
        obj = os.path.join(OBJDIR, unit + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_mtime):
            continue
        cmd = [NVCC] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for unit, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {unit} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(OUT):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
