"""TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/libstenos_oracle.so`` (built from ``oracle/stenos_oracle.c`` by
``make -C oracle oracle``): the plain-C restatement of the reference's level-1 path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libstenos_oracle.so")
NO_SHIFT = (1 << 64) - 1
ERR_BASE = (1 << 64) - 100

_lib = None


def build():
    src = os.path.join(_HERE, "stenos_oracle.c")
    if (not os.path.exists(SO_PATH)) or os.path.getmtime(SO_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO_PATH)
        sz, vp, ci = C.c_size_t, C.c_void_p, C.c_int
        L.so_shuffle.argtypes = [sz, sz, vp, vp]
        L.so_unshuffle.argtypes = [sz, sz, vp, vp]
        L.so_delta.argtypes = [vp, vp, sz]
        L.so_delta_inv.argtypes = [vp, vp, sz]
        for f in (L.so_shuffle, L.so_unshuffle, L.so_delta, L.so_delta_inv):
            f.restype = None
        L.so_block_compress.argtypes = [vp, sz, sz, vp, sz]
        L.so_block_compress.restype = sz
        L.so_block_decompress.argtypes = [vp, sz, sz, sz, vp]
        L.so_block_decompress.restype = sz
        L.so_compress_superblock.argtypes = [vp, sz, sz, vp, sz, ci]
        L.so_compress_superblock.restype = sz
        L.so_compress.argtypes = [vp, sz, sz, vp, sz, ci, sz]
        L.so_compress.restype = sz
        L.so_decompress.argtypes = [vp, sz, sz, vp, sz]
        L.so_decompress.restype = sz
        L.so_bound.argtypes = [sz]
        L.so_bound.restype = sz
        L.so_default_superblock.argtypes = [sz]
        L.so_default_superblock.restype = sz
        L.so_frame_index.argtypes = [vp, sz, sz, vp, sz]
        L.so_frame_index.restype = sz
        _lib = L
    return _lib


def has_error(r):
    return r >= ERR_BASE


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_u8(buf):
    a = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf)
    return a.view(np.uint8).reshape(-1)


def bound(n):
    return lib().so_bound(n)


def compress(buf, bytesoftype, level=1, dst_size=None, block_shift=None):
    L = lib()
    src = _as_u8(buf)
    if dst_size is None:
        dst_size = L.so_bound(src.size)
    dst = np.empty(max(dst_size, 1), dtype=np.uint8)
    r = L.so_compress(_ptr(src), bytesoftype, src.size, _ptr(dst), dst_size, level, NO_SHIFT if block_shift is None else block_shift)
    if has_error(r):
        raise RuntimeError("oracle so_compress error %d" % (r - (1 << 64)))
    return dst[:r].tobytes()


def decompress(cbuf, bytesoftype, out_bytes):
    L = lib()
    src = _as_u8(cbuf)
    dst = np.empty(max(out_bytes, 1), dtype=np.uint8)
    r = L.so_decompress(_ptr(src), bytesoftype, src.size, _ptr(dst), out_bytes)
    if has_error(r):
        raise RuntimeError("oracle so_decompress error %d" % (r - (1 << 64)))
    return dst[:r].tobytes()


def compress_superblock(buf, bytesoftype, level=1, room=None):
    L = lib()
    src = _as_u8(buf)
    if room is None:
        room = src.size + 4096
    dst = np.empty(room, dtype=np.uint8)
    r = L.so_compress_superblock(_ptr(src), bytesoftype, src.size, _ptr(dst), room, level)
    if has_error(r):
        raise RuntimeError("oracle so_compress_superblock error %d" % (r - (1 << 64)))
    return dst[:r].tobytes()


def block_decompress(payload, bytesoftype, out_bytes):
    L = lib()
    src = _as_u8(payload)
    dst = np.empty(max(out_bytes, 1), dtype=np.uint8)
    r = L.so_block_decompress(_ptr(src), src.size, bytesoftype, out_bytes, _ptr(dst))
    if has_error(r):
        raise RuntimeError("oracle so_block_decompress error %d" % (r - (1 << 64)))
    return dst[:out_bytes].tobytes(), r


def frame_index(cbuf, bytesoftype):
    L = lib()
    src = _as_u8(cbuf)
    cap = src.size // 4 + 4
    out = np.zeros(cap, dtype=np.uint64)
    r = L.so_frame_index(_ptr(src), bytesoftype, src.size, _ptr(out), cap)
    if has_error(r):
        raise RuntimeError("oracle so_frame_index error %d" % (r - (1 << 64)))
    return out[: r + 1].copy()


def shuffle(buf, bytesoftype):
    src = _as_u8(buf)
    dst = np.empty_like(src)
    lib().so_shuffle(bytesoftype, src.size, _ptr(src), _ptr(dst))
    return dst.tobytes()


def unshuffle(buf, bytesoftype):
    src = _as_u8(buf)
    dst = np.empty_like(src)
    lib().so_unshuffle(bytesoftype, src.size, _ptr(src), _ptr(dst))
    return dst.tobytes()


def delta(buf):
    src = _as_u8(buf)
    dst = np.empty_like(src)
    lib().so_delta(_ptr(src), _ptr(dst), src.size)
    return dst.tobytes()


def delta_inv(buf):
    src = _as_u8(buf)
    dst = np.empty_like(src)
    lib().so_delta_inv(_ptr(src), _ptr(dst), src.size)
    return dst.tobytes()
