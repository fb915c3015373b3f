"""Build libphaserot_cuda.so in-tree with nvcc for sm_100a.

The library is the product: hand-written CUDA kernels (csrc/kernels.cuh) behind
the C ABI of include/phaserot_cuda.h.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libphaserot_cuda.so")
HOST_DIR = os.path.join(HERE, "host")
BIN_DIR = os.path.join(HERE, "bin")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
    "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "phaserot_cuda.h")]
    if not force and not _newer(LIB, srcs):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "phaserot_cuda.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libphaserot_cuda.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


def build_host(force=False):
    """Host programs above the C ABI: the phase-rotate CLI and the LV2 plugin."""
    if not os.path.isdir(HOST_DIR):
        return []
    mk = os.path.join(HOST_DIR, "Makefile")
    if not os.path.exists(mk):
        return []
    args = ["make", "-s", "-C", HOST_DIR] + (["-B"] if force else [])
    subprocess.run(args, check=True)
    return [os.path.join(BIN_DIR, f) for f in sorted(os.listdir(BIN_DIR))] if os.path.isdir(BIN_DIR) else []


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    for p in build_host(force="--force" in sys.argv):
        print(p)
