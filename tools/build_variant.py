#!/usr/bin/env python
"""Experiment helper: builds the native library with extra nvcc flags into build/variants/<name>.so
(git-ignored, shipped by gpurun).  tools/time_parts.py picks a variant through STENOS_B200_LIB.

    python tools/build_variant.py nt768 -DSTREAM_THREADS_4=768
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stenos_b200 import build as b

name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build", "variants")
os.makedirs(out, exist_ok=True)
lib = os.path.join(out, name + ".so")
cmd = [b.nvcc()] + b.NVCC_FLAGS + extra + [os.path.join(b.CSRC, s) for s in b.SOURCES] + ["-o", lib, "-ldl"]
subprocess.check_call(cmd)
print("built", lib)
