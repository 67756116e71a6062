"""Builds fba-pomdp_b200/libfba_b200.so (CUDA kernels + the C ABI) for sm_100a with nvcc, in-tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "fba_capi.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "fba_kernels.cuh"), os.path.join(HERE, "csrc", "fba_device.cuh"),
        os.path.join(HERE, "..", "include", "fba_pomdp_b200.h")]
LIB = os.path.join(HERE, "libfba_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # replay parity: the reference runs on x86-64 without FMA, so never contract a*b+c
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-shared",
]


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
