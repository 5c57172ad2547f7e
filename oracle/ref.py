"""TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/_ref/libstenos_ref.so`` -- the untouched reference library built by
``oracle/Makefile`` (``make ref``) from ``/root/reference`` (public C API: ``stenos/stenos.h:115-301``;
filters: ``stenos/internal/shuffle.h:33,45`` and ``stenos/internal/delta.h:33,38``).

The prebuilt ``.so`` travels to the GPU box with gpurun; ``/root/reference`` itself does not.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "libstenos_ref.so")

_lib = None


def available():
    return os.path.exists(SO_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libstenos_ref.so missing: run `make -C oracle ref`")
        L = C.CDLL(SO_PATH)
        sz, vp, ci = C.c_size_t, C.c_void_p, C.c_int
        L.stenos_make_context.restype = vp
        L.stenos_destroy_context.argtypes = [vp]
        L.stenos_set_level.argtypes = [vp, ci]
        L.stenos_set_level.restype = sz
        L.stenos_set_threads.argtypes = [vp, ci]
        L.stenos_set_threads.restype = sz
        L.stenos_set_block_size.argtypes = [vp, sz]
        L.stenos_set_block_size.restype = sz
        L.stenos_bound.argtypes = [sz]
        L.stenos_bound.restype = sz
        for n in ("stenos_compress_generic", "stenos_decompress_generic"):
            f = getattr(L, n)
            f.argtypes = [vp, vp, sz, sz, vp, sz]
            f.restype = sz
        L.stenos_compress.argtypes = [vp, sz, sz, vp, sz, ci]
        L.stenos_compress.restype = sz
        L.stenos_decompress.argtypes = [vp, sz, sz, vp, sz]
        L.stenos_decompress.restype = sz
        for n in ("stenos_private_compress_block", "stenos_private_decompress_block"):
            f = getattr(L, n)
            f.argtypes = [vp, vp, sz, sz, sz, vp, sz]
            f.restype = sz
        L.stenos_private_create_compression_header.argtypes = [sz, sz, vp, sz]
        L.stenos_private_create_compression_header.restype = sz
        # C++ filter entry points (mangled), stenos::shuffle / unshuffle / delta / delta_inv
        for n in ("_ZN6stenos7shuffleEmmPKhPh", "_ZN6stenos9unshuffleEmmPKhPh"):
            f = getattr(L, n)
            f.argtypes = [sz, sz, vp, vp]
            f.restype = None
        for n in ("_ZN6stenos5deltaEPKvPvm", "_ZN6stenos9delta_invEPKvPvm"):
            f = getattr(L, n)
            f.argtypes = [vp, vp, sz]
            f.restype = None
        _lib = L
    return _lib


ERR_BASE = (1 << 64) - 100


def has_error(r):
    return r >= ERR_BASE


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_u8(buf):
    a = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf)
    return a.view(np.uint8).reshape(-1)


def bound(nbytes):
    return lib().stenos_bound(nbytes)


def compress(buf, bytesoftype, level=1, threads=1, dst_size=None, block_shift=None):
    """stenos_compress_generic on a fresh context. Returns bytes, raises on error."""
    L = lib()
    src = _as_u8(buf)
    n = src.size
    if dst_size is None:
        dst_size = L.stenos_bound(n)
    dst = np.empty(max(dst_size, 1), dtype=np.uint8)
    ctx = L.stenos_make_context()
    try:
        L.stenos_set_level(ctx, level)
        L.stenos_set_threads(ctx, threads)
        if block_shift is not None:
            L.stenos_set_block_size(ctx, block_shift)
        r = L.stenos_compress_generic(ctx, _ptr(src), bytesoftype, n, _ptr(dst), dst_size)
    finally:
        L.stenos_destroy_context(ctx)
    if has_error(r):
        raise RuntimeError("reference stenos_compress_generic error %d" % (r - (1 << 64)))
    return dst[:r].tobytes()


def decompress(cbuf, bytesoftype, out_bytes, threads=1):
    L = lib()
    src = _as_u8(cbuf)
    dst = np.empty(max(out_bytes, 1), dtype=np.uint8)
    ctx = L.stenos_make_context()
    try:
        L.stenos_set_threads(ctx, threads)
        r = L.stenos_decompress_generic(ctx, _ptr(src), bytesoftype, src.size, _ptr(dst), out_bytes)
    finally:
        L.stenos_destroy_context(ctx)
    if has_error(r):
        raise RuntimeError("reference stenos_decompress_generic error %d" % (r - (1 << 64)))
    return dst[:r].tobytes()


def compress_superblock(buf, bytesoftype, level=1, room=None):
    """stenos_private_compress_block: one superblock -> [code][csize:3][payload]."""
    L = lib()
    src = _as_u8(buf)
    n = src.size
    if room is None:
        room = n + 4096
    dst = np.empty(room, dtype=np.uint8)
    ctx = L.stenos_make_context()
    try:
        L.stenos_set_level(ctx, level)
        r = L.stenos_private_compress_block(ctx, _ptr(src), bytesoftype, max(n, 1), n, _ptr(dst), room)
    finally:
        L.stenos_destroy_context(ctx)
    if has_error(r):
        raise RuntimeError("reference stenos_private_compress_block error %d" % (r - (1 << 64)))
    return dst[:r].tobytes()


def decompress_superblock(cbuf, bytesoftype, out_bytes):
    L = lib()
    src = _as_u8(cbuf)
    dst = np.empty(max(out_bytes, 1), dtype=np.uint8)
    ctx = L.stenos_make_context()
    try:
        r = L.stenos_private_decompress_block(ctx, _ptr(src), bytesoftype, max(out_bytes, 1), src.size, _ptr(dst), out_bytes)
    finally:
        L.stenos_destroy_context(ctx)
    if has_error(r):
        raise RuntimeError("reference stenos_private_decompress_block error %d" % (r - (1 << 64)))
    return dst[:out_bytes].tobytes()


def shuffle(buf, bytesoftype):
    src = _as_u8(buf)
    dst = np.empty_like(src)
    lib()._ZN6stenos7shuffleEmmPKhPh(bytesoftype, src.size, _ptr(src), _ptr(dst))
    return dst.tobytes()


def unshuffle(buf, bytesoftype):
    src = _as_u8(buf)
    dst = np.empty_like(src)
    lib()._ZN6stenos9unshuffleEmmPKhPh(bytesoftype, src.size, _ptr(src), _ptr(dst))
    return dst.tobytes()


def delta(buf):
    src = _as_u8(buf)
    dst = np.empty_like(src)
    lib()._ZN6stenos5deltaEPKvPvm(_ptr(src), _ptr(dst), src.size)
    return dst.tobytes()


def delta_inv(buf):
    src = _as_u8(buf)
    dst = np.empty_like(src)
    lib()._ZN6stenos9delta_invEPKvPvm(_ptr(src), _ptr(dst), src.size)
    return dst.tobytes()
