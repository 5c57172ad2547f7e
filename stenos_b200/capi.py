"""ctypes binding of the C ABI in include/stenos_b200.h (libstenos_b200.so, built by nvcc).

This module is plumbing: it loads the native library and exposes its entry points with numpy /
raw-pointer arguments.  There is deliberately no CPU fallback: if the CUDA library is missing the
import of the product path fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstenos_b200.so")

NO_BLOCK_SHIFT = (1 << 64) - 1
ERR_BASE = (1 << 64) - 100
ERRORS = {
    (1 << 64) - 1: "UNDEFINED", (1 << 64) - 2: "SRC_OVERFLOW", (1 << 64) - 3: "ALLOC", (1 << 64) - 4: "INVALID_INPUT",
    (1 << 64) - 5: "INVALID_INSTRUCTION_SET", (1 << 64) - 6: "DST_OVERFLOW", (1 << 64) - 7: "INVALID_BYTESOFTYPE",
    (1 << 64) - 8: "ZSTD_INTERNAL", (1 << 64) - 9: "INVALID_PARAMETER",
}

# every symbol include/stenos_b200.h declares: name -> (restype, argtypes)
_sz, _vp, _ci, _u64 = C.c_size_t, C.c_void_p, C.c_int, C.c_uint64
SIGNATURES = {
    "stenos_make_context": (_vp, []),
    "stenos_destroy_context": (None, [_vp]),
    "stenos_reset_context": (None, [_vp]),
    "stenos_set_level": (_sz, [_vp, _ci]),
    "stenos_set_threads": (_sz, [_vp, _ci]),
    "stenos_set_max_nanoseconds": (_sz, [_vp, _u64]),
    "stenos_set_block_size": (_sz, [_vp, _sz]),
    "stenos_memory_footprint": (_sz, [_vp]),
    "stenos_has_error": (_ci, [_sz]),
    "stenos_bound": (_sz, [_sz]),
    "stenos_compress_generic": (_sz, [_vp, _vp, _sz, _sz, _vp, _sz]),
    "stenos_decompress_generic": (_sz, [_vp, _vp, _sz, _sz, _vp, _sz]),
    "stenos_compress": (_sz, [_vp, _sz, _sz, _vp, _sz, _ci]),
    "stenos_decompress": (_sz, [_vp, _sz, _sz, _vp, _sz]),
    "stenos_get_info": (_sz, [_vp, _sz, _sz, _vp]),
    "stenos_make_timer": (_vp, []),
    "stenos_destroy_timer": (None, [_vp]),
    "stenos_tick": (None, [_vp]),
    "stenos_tock": (_u64, [_vp]),
    "stenos_private_compress_block": (_sz, [_vp, _vp, _sz, _sz, _sz, _vp, _sz]),
    "stenos_private_decompress_block": (_sz, [_vp, _vp, _sz, _sz, _sz, _vp, _sz]),
    "stenos_private_block_size": (_sz, [_vp, _sz]),
    "stenos_private_block_csize": (_sz, [_vp]),
    "stenos_private_create_compression_header": (_sz, [_sz, _sz, _vp, _sz]),
    "stenos_set_device": (_sz, [_vp, _ci]),
    "stenos_set_stream": (_sz, [_vp, _vp]),
    "stenos_b200_compress_async": (_sz, [_vp, _vp, _sz, _sz, _vp, _sz, _vp, _vp]),
    "stenos_b200_decompress_async": (_sz, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _vp]),
    "stenos_b200_compress_segment_async": (_sz, [_vp, _vp, _sz, _sz, _vp, _sz, _vp, _vp]),
    "stenos_b200_decompress_range_async": (_sz, [_vp, _vp, _sz, _sz, _sz, _sz, _sz, _vp, _vp, _vp]),
    "stenos_b200_superblock_size": (_sz, [_vp, _sz, _sz]),
    "stenos_b200_frame_index_async": (_sz, [_vp, _vp, _sz, _sz, _vp, _sz, _vp]),
    "stenos_b200_index_accepted": (_ci, [_vp]),
    "stenos_b200_shuffle": (_sz, [_vp, _sz, _sz, _sz, _vp, _vp, _ci]),
    "stenos_b200_unshuffle": (_sz, [_vp, _sz, _sz, _sz, _vp, _vp, _ci]),
    "stenos_b200_delta": (_sz, [_vp, _sz, _sz, _vp, _vp]),
    "stenos_b200_delta_inv": (_sz, [_vp, _sz, _sz, _vp, _vp]),
    "stenos_b200_gather_decode_async": (_sz, [_vp, _vp, _sz, _sz, _sz, _sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    "stenos_b200_compress_strategy": (_sz, [_vp, _vp, _sz, _sz, _vp, _sz, C.c_int, C.c_int]),
    "stenos_b200_compress_buckets_async": (_sz, [_vp, _vp, _sz, _sz, _sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    "stenos_b200_synchronize": (_sz, [_vp]),
    "stenos_b200_test_occupy": (_sz, [_vp, _ci, C.c_uint, C.c_ulonglong]),
    "stenos_b200_kernel_launches": (C.c_ulonglong, []),
    "stenos_b200_build_target": (C.c_char_p, []),
}


class StenosError(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        self.name = ERRORS.get(code, "ERROR_%d" % (code - (1 << 64)))
        super().__init__("%s: STENOS_ERROR_%s" % (where, self.name))


def load(path=None):
    """Loads the native library and declares every prototype.  Raises if it is missing."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(
            "%s not found: the CUDA library must be built first (python -m stenos_b200.build); "
            "there is no CPU fallback for the device path" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        f.restype = res
        f.argtypes = args
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


def use_library(handle):
    """Tests only: route this module to another build of the same ABI (tests/emu)."""
    global _LIB
    _LIB = handle


def has_error(r):
    return r >= ERR_BASE


def check(r, where):
    if has_error(r):
        raise StenosError(r, where)
    return r


def ptr_of(obj):
    """Address of a numpy array, a torch tensor, an int, or None."""
    if obj is None:
        return None
    if isinstance(obj, int):
        return obj
    if isinstance(obj, np.ndarray):
        return obj.ctypes.data
    if hasattr(obj, "data_ptr"):
        return obj.data_ptr()
    raise TypeError("cannot take the address of %r" % type(obj))


def as_u8(buf):
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    return np.frombuffer(buf, dtype=np.uint8)
