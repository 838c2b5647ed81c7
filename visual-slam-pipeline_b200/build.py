"""Builds libvsm.so (the CUDA matching library + its C ABI) in-tree for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the
resulting lib/libvsm.so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libvsm.so")
SOURCES = ["vsm_api.cu"]
DEPS = ["vsm_api.cu", "vsm_group.inl", "vsm_tc.cuh", "vsm_tc2.cuh", "vsm_kernels.cuh", "vsm_common.cuh", os.path.join("..", "..", "include", "vsm.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_lib(force=False, verbose=False):
    if not (force or stale()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
