"""Whole-buffer partitioning of one stenos frame across the GPUs of a box.

Superblocks are independent in both directions (stenos.cpp:893-904, :1124-1143), so rank g takes a
contiguous range of whole superblocks, encodes it into a SEGMENT (superblocks back to back, no frame
header) in its own HBM, and only the segment byte lengths are exchanged (one 8-byte integer per rank,
all-gathered) to turn them into frame offsets.  No collective touches the data path.  An optional
gather of the segments to one rank exists for callers that want one contiguous stream.

One process per GPU; `torch.distributed` (nccl on GPUs, gloo in the CPU tests) is plumbing only.
"""
import numpy as np


def plan_partition(total_bytes, superblock_bytes, world_size):
    """Contiguous superblock ranges per rank: [(first_sb, n_sb, byte_offset, byte_count)]."""
    n_sb = (total_bytes + superblock_bytes - 1) // superblock_bytes
    out = []
    for g in range(world_size):
        lo = (g * n_sb) // world_size
        hi = ((g + 1) * n_sb) // world_size
        b0 = lo * superblock_bytes
        b1 = min(hi * superblock_bytes, total_bytes)
        out.append((lo, hi - lo, b0, max(b1 - b0, 0)))
    return out


def frame_header(total_bytes, superblock_bytes=None):
    """[shift][decompressed bytes:7] (+[superblock bytes:4] for custom sizes) -- stenos.cpp:862-874."""
    if superblock_bytes is None:
        return bytes([0]) + int(total_bytes).to_bytes(7, "little")
    return bytes([255]) + int(total_bytes).to_bytes(7, "little") + int(superblock_bytes).to_bytes(4, "little")


def exchange_segment_sizes(my_size, dist, device):
    """All-gathers the segment lengths; returns (sizes[world], my exclusive prefix)."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    mine = my_size if hasattr(my_size, "device") else torch.tensor([int(my_size)], dtype=torch.int64, device=device)
    mine = mine.reshape(1).to(torch.int64)
    got = torch.empty(world, dtype=torch.int64, device=mine.device)
    try:
        dist.all_gather_into_tensor(got, mine)
    except (RuntimeError, AttributeError):
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        got = torch.cat(parts)
    sizes = got.cpu().numpy().astype(np.int64)
    return sizes, int(sizes[:rank].sum())


class SegmentCodec:
    """Rank-local encoder / decoder of one segment of a partitioned frame (device resident buffers)."""

    def __init__(self, ctx, bytesoftype, frame_bytes):
        self.ctx = ctx
        self.T = bytesoftype
        self.frame_bytes = frame_bytes
        self.sb = ctx.superblock_size(bytesoftype, frame_bytes)

    def capacity(self, seg_bytes):
        n_sb = (seg_bytes + self.sb - 1) // self.sb
        return seg_bytes + 4 * n_sb + 16

    def compress_async(self, d_src, seg_bytes, d_dst, dst_size, d_result, d_sb_offsets=None):
        self.ctx.superblock_size(self.T, self.frame_bytes)  # the segment uses the FRAME's superblock size
        return self.ctx.compress_segment_async(d_src, self.T, seg_bytes, d_dst, dst_size, d_result, d_sb_offsets)

    def decompress_async(self, d_seg, seg_csize, seg_bytes, d_sb_offsets, d_dst, d_result):
        self.ctx.superblock_size(self.T, self.frame_bytes)
        n_sb = (seg_bytes + self.sb - 1) // self.sb
        return self.ctx.decompress_range_async(d_seg, seg_csize, self.T, seg_bytes, 0, n_sb, d_sb_offsets, d_dst, d_result)
