"""Build libdgb200.so (hand-written sm_100a CUDA + the C ABI of include/dgb200.h) in-tree.

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the snapshot."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdgb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--threads", "0", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    # fp64 parity: no fast-math; fma contraction stays on (explicit fma() is used anyway)
    "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(HERE), "include", "dgb200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # DGB_NVCC_EXTRA / DGB_LIB_OUT: diagnostic builds next to the product library (e.g. -DDGB_CHAIN_TRACE)
    extra = os.environ.get("DGB_NVCC_EXTRA", "").split()
    out = os.environ.get("DGB_LIB_OUT", LIB)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdgb200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
