"""Builds csrc/*.cu into liblbic_b200.so (in this directory) with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  Usage: python build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
SO = os.path.join(HERE, "liblbic_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
SOURCES = ["api.cu", "gemm_tc.cu", "gemm_ws.cu", "gemm_wave.cu", "gemm_simt.cu", "kernels_misc.cu", "tables.cu", "rans.cu", "metrics.cu"]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "lbic.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _newer(src, obj) or hdr_time > os.path.getmtime(obj):
            cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=6) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if force or jobs or not os.path.exists(SO):
        run([NVCC, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
