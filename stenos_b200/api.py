"""Host-side mirror of the reference's interface for the level-1 path (same names, argument meaning
and error behaviour as stenos/stenos.h), on top of the C ABI.  numpy arrays / bytes are HOST
buffers; torch CUDA tensors (or raw integer addresses) are DEVICE buffers.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import StenosError, as_u8, check, has_error, ptr_of  # noqa: F401


def bound(nbytes):
    """stenos_bound (stenos/stenos.h:185)."""
    return capi.lib().stenos_bound(nbytes)


class Context:
    """stenos_context (stenos/stenos.h:103-173) plus the device selector / stream knobs."""

    def __init__(self, level=1, device=None, stream=None, block_shift=None):
        self._lib = capi.lib()
        self._h = self._lib.stenos_make_context()
        if not self._h:
            raise MemoryError("stenos_make_context")
        self.set_level(level)
        if device is not None:
            self.set_device(device)
        if stream is not None:
            self.set_stream(stream)
        if block_shift is not None:
            self.set_block_size(block_shift)

    def close(self):
        if self._h:
            self._lib.stenos_destroy_context(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # --- knobs
    def set_level(self, level):
        return check(self._lib.stenos_set_level(self._h, level), "stenos_set_level")

    def set_threads(self, threads):
        return check(self._lib.stenos_set_threads(self._h, threads), "stenos_set_threads")

    def set_max_nanoseconds(self, ns):
        return check(self._lib.stenos_set_max_nanoseconds(self._h, ns), "stenos_set_max_nanoseconds")

    def set_block_size(self, shift):
        return check(self._lib.stenos_set_block_size(self._h, capi.NO_BLOCK_SHIFT if shift is None else shift), "stenos_set_block_size")

    def set_device(self, device):
        return check(self._lib.stenos_set_device(self._h, device), "stenos_set_device")

    def set_stream(self, stream):
        """stream: a cudaStream_t as int, or a torch.cuda.Stream."""
        h = getattr(stream, "cuda_stream", stream)
        return check(self._lib.stenos_set_stream(self._h, h), "stenos_set_stream")

    def reset(self):
        self._lib.stenos_reset_context(self._h)

    def memory_footprint(self):
        return self._lib.stenos_memory_footprint(self._h)

    def synchronize(self):
        return check(self._lib.stenos_b200_synchronize(self._h), "stenos_b200_synchronize")

    def superblock_size(self, bytesoftype, nbytes):
        return check(self._lib.stenos_b200_superblock_size(self._h, bytesoftype, nbytes), "stenos_b200_superblock_size")

    # --- raw calls (addresses; host or device)
    def compress_raw(self, src, bytesoftype, nbytes, dst, dst_size):
        return self._lib.stenos_compress_generic(self._h, ptr_of(src), bytesoftype, nbytes, ptr_of(dst), dst_size)

    def decompress_raw(self, src, bytesoftype, nbytes, dst, dst_size):
        return self._lib.stenos_decompress_generic(self._h, ptr_of(src), bytesoftype, nbytes, ptr_of(dst), dst_size)

    # --- host convenience
    def compress(self, buf, bytesoftype, dst_size=None):
        """stenos_compress_generic on a host buffer; returns bytes."""
        src = as_u8(buf)
        if dst_size is None:
            dst_size = bound(src.size)
        dst = np.empty(max(dst_size, 1), dtype=np.uint8)
        r = check(self.compress_raw(src, bytesoftype, src.size, dst, dst_size), "stenos_compress_generic")
        return dst[:r].tobytes()

    def decompress(self, cbuf, bytesoftype, out_bytes):
        src = as_u8(cbuf)
        dst = np.empty(max(out_bytes, 1), dtype=np.uint8)
        r = check(self.decompress_raw(src, bytesoftype, src.size, dst, out_bytes), "stenos_decompress_generic")
        return dst[:r].tobytes()

    def compress_block(self, buf, bytesoftype, super_block_size=None, room=None):
        """stenos_private_compress_block (the stenos::cvector bucket codec)."""
        src = as_u8(buf)
        if super_block_size is None:
            super_block_size = max(src.size, 1)
        if room is None:
            room = src.size + 4096
        dst = np.empty(room, dtype=np.uint8)
        r = check(self._lib.stenos_private_compress_block(self._h, ptr_of(src), bytesoftype, super_block_size, src.size, ptr_of(dst), room),
                  "stenos_private_compress_block")
        return dst[:r].tobytes()

    def decompress_block(self, cbuf, bytesoftype, out_bytes, super_block_size=None):
        src = as_u8(cbuf)
        dst = np.empty(max(out_bytes, 1), dtype=np.uint8)
        r = check(self._lib.stenos_private_decompress_block(self._h, ptr_of(src), bytesoftype, super_block_size or max(out_bytes, 1), src.size, ptr_of(dst), out_bytes),
                  "stenos_private_decompress_block")
        return dst[:r].tobytes()

    # --- filters (host or device buffers decided by the pointer)
    def shuffle(self, buf, bytesoftype, chunk=0, with_delta=False):
        src = as_u8(buf)
        dst = np.empty_like(src)
        check(self._lib.stenos_b200_shuffle(self._h, bytesoftype, src.size, chunk, ptr_of(src), ptr_of(dst), int(with_delta)), "stenos_b200_shuffle")
        return dst.tobytes()

    def unshuffle(self, buf, bytesoftype, chunk=0, with_delta=False):
        src = as_u8(buf)
        dst = np.empty_like(src)
        check(self._lib.stenos_b200_unshuffle(self._h, bytesoftype, src.size, chunk, ptr_of(src), ptr_of(dst), int(with_delta)), "stenos_b200_unshuffle")
        return dst.tobytes()

    def delta(self, buf, chunk=0):
        src = as_u8(buf)
        dst = np.empty_like(src)
        check(self._lib.stenos_b200_delta(self._h, src.size, chunk, ptr_of(src), ptr_of(dst)), "stenos_b200_delta")
        return dst.tobytes()

    def delta_inv(self, buf, chunk=0):
        src = as_u8(buf)
        dst = np.empty_like(src)
        check(self._lib.stenos_b200_delta_inv(self._h, src.size, chunk, ptr_of(src), ptr_of(dst)), "stenos_b200_delta_inv")
        return dst.tobytes()

    # --- device resident, asynchronous (addresses or torch tensors)
    def compress_async(self, d_src, bytesoftype, nbytes, d_dst, dst_size, d_result, d_sb_offsets=None):
        return check(self._lib.stenos_b200_compress_async(self._h, ptr_of(d_src), bytesoftype, nbytes, ptr_of(d_dst), dst_size, ptr_of(d_result), ptr_of(d_sb_offsets)),
                     "stenos_b200_compress_async")

    def decompress_async(self, d_src, bytesoftype, nbytes, d_dst, dst_size, out_bytes, d_result, d_sb_offsets=None):
        return check(self._lib.stenos_b200_decompress_async(self._h, ptr_of(d_src), bytesoftype, nbytes, ptr_of(d_dst), dst_size, out_bytes, ptr_of(d_result),
                                                             ptr_of(d_sb_offsets)), "stenos_b200_decompress_async")

    def compress_segment_async(self, d_src, bytesoftype, seg_bytes, d_dst, dst_size, d_result, d_sb_offsets=None):
        return check(self._lib.stenos_b200_compress_segment_async(self._h, ptr_of(d_src), bytesoftype, seg_bytes, ptr_of(d_dst), dst_size, ptr_of(d_result),
                                                                   ptr_of(d_sb_offsets)), "stenos_b200_compress_segment_async")

    def decompress_range_async(self, d_frame, frame_bytes, bytesoftype, out_bytes, first_sb, n_sb, d_sb_offsets, d_dst, d_result):
        return check(self._lib.stenos_b200_decompress_range_async(self._h, ptr_of(d_frame), frame_bytes, bytesoftype, out_bytes, first_sb, n_sb, ptr_of(d_sb_offsets),
                                                                   ptr_of(d_dst), ptr_of(d_result)), "stenos_b200_decompress_range_async")

    def frame_index_async(self, d_frame, frame_bytes, bytesoftype, d_sb_offsets, capacity, d_result):
        return check(self._lib.stenos_b200_frame_index_async(self._h, ptr_of(d_frame), frame_bytes, bytesoftype, ptr_of(d_sb_offsets), capacity, ptr_of(d_result)),
                     "stenos_b200_frame_index_async")

    def index_accepted(self):
        """1: the last parallel frame index was accepted, 0: it fell back to the serial walk, -1: none ran."""
        return int(self._lib.stenos_b200_index_accepted(self._h))

    def compress_strategy(self, buf, bytesoftype, level, strategy, dst_size=None):
        """Level >= 2 with a forced strategy: 3 = Zstd(shuffle), 4 = Zstd(delta(shuffle)) (stenos.cpp:617-656)."""
        src = as_u8(buf)
        if dst_size is None:
            dst_size = bound(src.size)
        dst = np.empty(max(dst_size, 1), dtype=np.uint8)
        r = check(self._lib.stenos_b200_compress_strategy(self._h, ptr_of(src), bytesoftype, src.size, ptr_of(dst), dst_size, level, strategy), "stenos_b200_compress_strategy")
        return dst[:r].tobytes()

    def compress_buckets_async(self, d_src, bytesoftype, bucket_bytes, total_bytes, d_ids, n, d_slots, slot_stride, d_sizes, d_result):
        """n cvector buckets in one launch (cvector.hpp:1394-1420): bare superblocks in fixed slots + their sizes."""
        return check(self._lib.stenos_b200_compress_buckets_async(self._h, ptr_of(d_src), bytesoftype, bucket_bytes, total_bytes, ptr_of(d_ids), n, ptr_of(d_slots), slot_stride,
                                                                   ptr_of(d_sizes), ptr_of(d_result)), "stenos_b200_compress_buckets_async")

    def gather_decode_async(self, d_frame, frame_bytes, bytesoftype, bucket_bytes, out_bytes, d_sb_offsets, n_buckets, d_ids, n, d_dst, d_result):
        return check(self._lib.stenos_b200_gather_decode_async(self._h, ptr_of(d_frame), frame_bytes, bytesoftype, bucket_bytes, out_bytes, ptr_of(d_sb_offsets), n_buckets,
                                                                ptr_of(d_ids), n, ptr_of(d_dst), ptr_of(d_result)), "stenos_b200_gather_decode_async")


def compress(buf, bytesoftype, level=1, dst_size=None):
    """stenos_compress (stenos/stenos.h:224) on a host buffer."""
    src = as_u8(buf)
    if dst_size is None:
        dst_size = bound(src.size)
    dst = np.empty(max(dst_size, 1), dtype=np.uint8)
    r = check(capi.lib().stenos_compress(ptr_of(src), bytesoftype, src.size, ptr_of(dst), dst_size, level), "stenos_compress")
    return dst[:r].tobytes()


def decompress(cbuf, bytesoftype, out_bytes):
    """stenos_decompress (stenos/stenos.h:237) on a host buffer."""
    src = as_u8(cbuf)
    dst = np.empty(max(out_bytes, 1), dtype=np.uint8)
    r = check(capi.lib().stenos_decompress(ptr_of(src), bytesoftype, src.size, ptr_of(dst), out_bytes), "stenos_decompress")
    return dst[:r].tobytes()


def get_info(cbuf, bytesoftype):
    """stenos_get_info (stenos/stenos.h:256): (decompressed_size, superblock_size, header_bytes)."""
    src = as_u8(cbuf)

    class Info(C.Structure):
        _fields_ = [("decompressed_size", C.c_size_t), ("superblock_size", C.c_size_t)]

    info = Info()
    r = check(capi.lib().stenos_get_info(ptr_of(src), bytesoftype, src.size, C.addressof(info)), "stenos_get_info")
    return info.decompressed_size, info.superblock_size, r


def kernel_launches():
    return capi.lib().stenos_b200_kernel_launches()
