"""Builds the in-tree native library  stenos_b200/libstenos_b200.so  with nvcc for sm_100a.

    python -m stenos_b200.build            # build if sources are newer than the library
    python -m stenos_b200.build --force

The library is git-ignored (history stays source-only) but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libstenos_b200.so")
SOURCES = ["sb_api.cu"]
HEADERS = ["sb_common.cuh", "sb_encode.cuh", "sb_encode_rows.cuh", "sb_decode.cuh", "sb_decode_rows.cuh", "sb_kernels.cuh", "sb_stream.cuh", "sb_flow.cuh", "sb_filters.cuh", os.path.join("..", "..", "include", "stenos_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
    "-Xlinker", "-Bsymbolic-functions",
    "--shared", "-cudart", "static",
]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + list(extra) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB, "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else ())
    print("built", LIB)
