"""TEST INFRASTRUCTURE ONLY: builds and loads tests/emu/libstenos_b200_emu.so -- the product's CUDA
sources compiled by g++ against the SIMT emulator (tests/emu/cuda_emu.h) -- and routes
stenos_b200.capi to it, so kernels and host logic are exercised on machines without a GPU."""
import os
import subprocess

from stenos_b200 import capi

_EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_handle = None


def load():
    global _handle
    if _handle is None:
        subprocess.check_call(["make", "-s", "-C", _EMU_DIR])
        _handle = capi.load(os.path.join(_EMU_DIR, "libstenos_b200_emu.so"))
        assert _handle.stenos_b200_build_target() == b"emu"
    return _handle


def activate():
    capi.use_library(load())
