"""Builds libfaldoi_gpu.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python faldoi-ipol_b200/build.py [--force]

nvcc cross-compiles without a GPU.  -fmad=false is part of the numerical
contract (bit-comparable with the reference's non-FMA fp32 arithmetic);
division and sqrt stay IEEE (no --use_fast_math).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfaldoi_gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-shared", "-Xptxas", "-v",
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in sources() + [os.path.join(HERE, "..", "include", "faldoi_gpu.h")])


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    cmd = [NVCC] + FLAGS + ["-o", OUT, os.path.join(CSRC, "faldoi_gpu.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
