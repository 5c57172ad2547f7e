"""CPU-only: the product's CUDA sources compiled against the SIMT emulator (tests/emu) and checked
bit-for-bit against the oracle.  This is a unit test of kernel + host logic on a machine without a
GPU -- NOT a product code path (the product library is the nvcc build, see test_capi.py)."""
import numpy as np
import pytest

import dists
import emu_lib
from oracle import port, ref
from stenos_b200 import api, capi


@pytest.fixture(scope="module", autouse=True)
def emulated_library():
    prev = capi._LIB
    emu_lib.activate()
    yield
    capi.use_library(prev)


def raw_of(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def run(fn, *a, **k):
    try:
        return fn(*a, **k)
    except api.StenosError as e:
        return e.name
    except RuntimeError as e:
        return {"-6": "DST_OVERFLOW", "-2": "SRC_OVERFLOW", "-4": "INVALID_INPUT"}.get(str(e).split()[-1], str(e))


@pytest.mark.parametrize("T", [2, 4, 8, 3, 5, 6, 7])  # 3, 5, 6, 7: the generic kernels
def test_frames_match_oracle(T):
    for name in dists.names():
        for n in (256, 700, 33):
            raw = raw_of(dists.make(name, n, T, seed=5))
            want = port.compress(raw, T)
            assert api.compress(raw, T) == want, (name, n)
            assert api.decompress(want, T, raw.size) == raw.tobytes(), (name, n)


def test_multi_superblock_frame_and_index():
    for T, name in ((4, "ramp_noise16"), (8, "lz_then_noise"), (2, "smooth_sine")):
        n = (131072 * 2 + 5000) // T
        raw = raw_of(dists.make(name, n, T, seed=9))
        want = port.compress(raw, T)
        got = api.compress(raw, T)
        assert got == want
        assert api.decompress(got, T, raw.size) == raw.tobytes()


def test_exact_multiple_of_superblock_roundtrip():
    raw = raw_of(dists.make("sorted", 32768, 4))
    c = api.compress(raw, 4)
    assert c == port.compress(raw, 4)
    assert api.decompress(c, 4, raw.size) == raw.tobytes()  # the reference decoder fails here (appendix C1)


def test_tiny_and_empty_inputs():
    ctx = api.Context()
    assert ctx.compress(b"", 4) == port.compress(np.zeros(0, np.uint8), 4)
    assert ctx.decompress(port.compress(np.zeros(0, np.uint8), 4), 4, 0) == b""
    for n in (1, 7, 31, 32, 33, 63):  # < 128 bytes: Zstd / COPY superblock through libzstd
        raw = raw_of(dists.make("ramp_noise4", n, 4, seed=n))[: 4 * n]
        want = port.compress(raw, 4)
        assert ctx.compress(raw, 4) == want
        assert ctx.decompress(want, 4, raw.size) == raw.tobytes()
    raw = raw_of(dists.make("sorted", 8192 + 10, 4))  # full superblock + 40 byte tail
    want = port.compress(raw, 4)
    assert ctx.compress(raw, 4) == want
    assert ctx.decompress(want, 4, raw.size) == raw.tobytes()


def test_level0_and_unsupported_parameters():
    raw = raw_of(dists.make("sorted", 40000, 4))
    assert api.compress(raw, 4, level=0) == port.compress(raw, 4, level=0)
    assert api.decompress(port.compress(raw, 4, level=0), 4, raw.size) == raw.tobytes()
    assert run(api.compress, raw, 4, level=2) == "INVALID_PARAMETER"  # Zstd levels: no CPU fallback
    assert run(api.compress, raw[:39996], 9) == "INVALID_PARAMETER"  # element sizes above 8 (and 1) are not built
    assert run(api.compress, raw, 0) == "INVALID_BYTESOFTYPE"
    ctx = api.Context()
    ctx.set_max_nanoseconds(1000)
    assert run(ctx.compress, raw, 4) == "INVALID_PARAMETER"


def test_room_dependent_decisions():
    ctx = api.Context()
    for T in (4, 8, 3):
        for name in ("random", "lz_pairs", "repeat_period7", "sorted", "mostly_random_some_repeats"):
            raw = raw_of(dists.make(name, 256, T, seed=1))
            for room in (T * 256 + 16, T * 256 + 4, T * 256 + 64, T * 256 + 300):
                w = run(port.compress_superblock, raw, T, room=room)
                assert run(ctx.compress_block, raw, T, room=room) == w, (T, name, room)
                if isinstance(w, bytes):
                    assert ctx.decompress_block(w, T, raw.size) == raw.tobytes()
            raw = raw_of(dists.make(name, 3000, T, seed=2))
            full = len(port.compress(raw, T))
            for ds in (full + 8, full, full - 1, raw.size, 20, 7):
                assert run(ctx.compress, raw, T, dst_size=ds) == run(port.compress, raw, T, dst_size=ds), (T, name, ds)


def test_cvector_style_frame():
    # custom superblock = one 256 element block per bucket, 12 byte header (cvector.hpp:3034-3093)
    ctx = api.Context(block_shift=0)
    raw = raw_of(dists.make("sparse_changes", 256 * 9 + 100, 4, seed=3))
    want = port.compress(raw, 4, block_shift=0)
    assert ctx.compress(raw, 4) == want
    assert api.decompress(want, 4, raw.size) == raw.tobytes()


def test_corrupt_input_is_rejected_not_overrun():
    raw = raw_of(dists.make("ramp_noise16", 3000, 4, seed=4))
    c = bytearray(port.compress(raw, 4))
    assert run(api.decompress, bytes(c[: len(c) // 2]), 4, raw.size) in ("SRC_OVERFLOW", "INVALID_INPUT")
    c2 = bytearray(c)
    c2[8] = 9  # unknown superblock code
    assert run(api.decompress, bytes(c2), 4, raw.size) == "INVALID_INPUT"
    assert run(api.decompress, bytes(c), 4, raw.size - 1) == "DST_OVERFLOW"
    rng = np.random.default_rng(0)
    for _ in range(20):
        c3 = bytearray(c)
        for k in rng.integers(12, len(c), 4):
            c3[k] ^= int(rng.integers(1, 256))
        r = run(api.decompress, bytes(c3), 4, raw.size)  # must return (anything) without crashing
        assert isinstance(r, (bytes, str))


def test_filters_match_oracle():
    ctx = api.Context()
    rng = np.random.default_rng(3)
    for T in (2, 4, 8):
        for n in (0, 1, 100, 2048 // T + 1, 5000, 16384):
            a = rng.integers(0, 256, n * T + (int(rng.integers(0, T)) if n else 0), dtype=np.uint8)
            for chunk in (0, 4096, 2080 if T == 2 else 16 * T * 33):  # the last: fused unshuffle+delta_inv, T=2 with mid-group stream starts
                step = chunk or max(a.size, 1)
                pieces = [a[i:i + step] for i in range(0, a.size, step)]
                sh = b"".join(port.shuffle(p, T) for p in pieces)
                shd = b"".join(port.delta(np.frombuffer(port.shuffle(p, T), dtype=np.uint8)) for p in pieces)
                assert ctx.shuffle(a, T, chunk) == sh
                assert ctx.shuffle(a, T, chunk, True) == shd
                assert ctx.unshuffle(np.frombuffer(sh, dtype=np.uint8), T, chunk) == a.tobytes()
                assert ctx.unshuffle(np.frombuffer(shd, dtype=np.uint8), T, chunk, True) == a.tobytes()
            assert ctx.delta(a) == port.delta(a)
            assert ctx.delta_inv(np.frombuffer(port.delta(a), dtype=np.uint8)) == a.tobytes()


def test_parallel_frame_index_resynchronises_or_falls_back():
    """Frames whose payload imitates [code][csize:3] headers: the parallel index must still be exact."""
    ctx = api.Context(block_shift=0)
    n_sb = 9000
    res = np.zeros(2, dtype=np.uint64)
    rng = np.random.default_rng(1)
    a = rng.integers(-2**31, 2**31, 256 * n_sb).astype(np.int32)
    a[: 256 * 3000: 3] = 7
    frames = [port.compress(a, 4, block_shift=0, dst_size=a.nbytes + 8 * n_sb + 64)]
    for fake in (1, 0x00001001):  # "01 00 00 00" = code 1, csize 0 ; code 1, csize 16 -> plausible multi-hop chains
        b = np.full(256 * n_sb, fake, dtype=np.int32)
        frames.append(port.compress(b, 4, level=0, block_shift=0, dst_size=b.nbytes + 8 * n_sb + 64))
    for i, frame in enumerate(frames):
        f = np.frombuffer(frame, dtype=np.uint8).copy()
        offs = np.zeros(n_sb + 1, dtype=np.uint64)
        l0 = api.kernel_launches()
        assert ctx.frame_index_async(f, len(frame), 4, offs, n_sb + 1, res) == n_sb
        ctx.synchronize()
        assert api.kernel_launches() - l0 == 3  # the parallel path (scan + merge + fill), not the serial kernel
        assert res[1] == 0 and np.array_equal(offs, port.frame_index(frame, 4))
        if i == 0:
            assert ctx.index_accepted() == 1  # an ordinary frame must not need the serial fallback


def _mixed(T, n_sb, seed, tail_elems):
    """n_sb superblocks that alternate between incompressible, LZ-friendly and well compressible data + a partial tail."""
    per = 131072 // T
    kinds = ["random", "ramp_noise16", "random", "lz_then_noise", "random", "random", "sparse_changes", "mostly_random_some_repeats", "const"]
    parts = [raw_of(dists.make(kinds[(i + seed) % len(kinds)], per, T, seed=seed + i)) for i in range(n_sb)]
    if tail_elems:
        parts.append(raw_of(dists.make("random" if seed & 1 else "ramp_noise16", tail_elems, T, seed=seed)))
    return np.concatenate(parts)


@pytest.mark.parametrize("T,n_sb,tail", [(4, 11, 255 * 4 // 4), (8, 7, 37), (2, 6, 0)])
def test_stream_encoder_ring_wrap_copy_superblocks_and_tail(T, n_sb, tail):
    """encode_stream_kernel: ample dst room -> every superblock goes through the barrier-free pipeline; runs of
    incompressible superblocks fill the shared-memory ring (wrap, waits for flushes), COPY decisions, partial tail."""
    raw = _mixed(T, n_sb, seed=T, tail_elems=tail)
    room = raw.size + 300000
    want = port.compress(raw, T, dst_size=room)
    ctx = api.Context()
    l0 = api.kernel_launches()
    got = ctx.compress(raw, T, dst_size=room)
    assert api.kernel_launches() - l0 == 1  # the stream kernel alone
    assert got == want
    assert api.decompress(got, T, raw.size) == raw.tobytes()


def test_stream_encoder_many_tiny_superblocks():
    """cvector-style frame (one 256-element block per superblock) with ample room: hundreds of superblocks per CTA,
    two tickets each -- exercises slot recycling, the global ticket counter and termination of idle warps."""
    ctx = api.Context(block_shift=0)
    raw = np.concatenate([raw_of(dists.make(n, 256 * 40, 4, seed=i)) for i, n in enumerate(("sparse_changes", "random", "lz_pairs", "const", "ramp_noise16"))]
                         + [raw_of(dists.make("sorted", 77, 4))])
    room = raw.size * 4 + 4096
    want = port.compress(raw, 4, block_shift=0, dst_size=room)
    l0 = api.kernel_launches()
    assert ctx.compress(raw, 4, dst_size=room) == want
    assert api.kernel_launches() - l0 == 1
    assert api.decompress(want, 4, raw.size) == raw.tobytes()


@pytest.mark.parametrize("T", [2, 4, 8])
def test_row_decoder_every_distribution(T):
    """decode_pairs_kernel's lane-per-row path only takes blocks that lie a worst-case block before the end of the
    buffer: put every distribution in front of an incompressible tail so that its blocks are decoded there, and
    check it against the barrier-free decoder's predecessor (STENOS_B200_LEGACY_DECODER) as well."""
    import os
    tail = raw_of(dists.make("random", 256 * 3, T, seed=77))
    for name in dists.names():
        raw = np.concatenate([raw_of(dists.make(name, 256 * 5, T, seed=11)), tail])
        c = port.compress(raw, T)
        assert api.decompress(c, T, raw.size) == raw.tobytes(), name
    os.environ["STENOS_B200_LEGACY_DECODER"] = "1"
    try:
        assert api.Context().decompress(c, T, raw.size) == raw.tobytes()
    finally:
        del os.environ["STENOS_B200_LEGACY_DECODER"]


def test_pipelined_host_path_matches_single_launch(monkeypatch):
    """stenos_compress_generic on host buffers: chunks of whole superblocks encoded as segments while later chunks
    are still being copied in; the concatenation must be the frame (incl. the < 128 byte Zstd tail)."""
    for T, n in ((4, (5 * 131072 + 4 * 256 * 3 + 100) // 4), (2, (4 * 131072 + 60) // 2), (8, (3 * 131072 + 8 * 300) // 8)):
        rng = np.random.default_rng(11)
        a = (np.arange(n) * 3 + rng.integers(0, 16, n)).astype({2: np.int16, 4: np.int32, 8: np.int64}[T])
        raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        want = port.compress(raw, T)
        monkeypatch.setenv("STENOS_B200_PIPELINE_CHUNK", str(131072))
        got = run(api.compress, raw, T)
        monkeypatch.setenv("STENOS_B200_PIPELINE_CHUNK", "0")
        single = run(api.compress, raw, T)
        assert got == want and single == want
        assert run(api.decompress, got, T, raw.size) == raw.tobytes()


def test_flow_encoder_spill_path_with_minimum_rings(monkeypatch):
    """The fast encoder (sb_flow.cuh) with its staging rings at the smallest legal size: pieces whose superblock is not
    placed yet move to HBM spill slots all the time.  Streams must not change."""
    monkeypatch.setenv("STENOS_B200_FLOW_RING", "1")
    for T, name in ((4, "mostly_random_some_repeats"), (8, "lz_then_noise"), (2, "random"), (4, "ramp_noise200")):
        n = (131072 * 6 + 5000) // T
        raw = raw_of(dists.make(name, n, T, seed=11))
        want = port.compress(raw, T)
        # room beyond stenos_bound: every superblock goes through the fast encoder
        assert api.Context().compress(raw, T, dst_size=api.bound(raw.size) + 900000) == want, (T, name)


def test_pipelined_host_decompress(monkeypatch):
    """stenos_decompress_generic on host buffers in chunks (H2D, decode, D2H overlapped): small chunks force the path."""
    monkeypatch.setenv("STENOS_B200_PIPELINE_CHUNK", "262144")
    for T, name in ((4, "ramp_noise16"), (2, "smooth_sine"), (8, "lz_then_noise"), (4, "random")):
        for nb in (131072 * 9 + 64 * T, 131072 * 5 + 300 * T + 3, 131072 * 8):
            raw = raw_of(dists.make(name, nb // T, T, seed=3))
            c = port.compress(raw, T)
            assert api.compress(raw, T) == c
            assert api.decompress(c, T, raw.size) == raw.tobytes(), (T, name, nb)
    raw = raw_of(dists.make("ramp_noise16", 131072 * 3, 4, seed=1))
    c = port.compress(raw, 4)
    for cut in (len(c) - 5, len(c) // 2):
        assert run(api.decompress, c[:cut], 4, raw.size) == "INVALID_INPUT"


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the compiled reference) writes the level >= 2 frames")
def test_hybrid_decoder_reads_the_reference_frames_of_every_level():
    """Frames the REFERENCE writes at levels 2..9 (superblock codes 2..5: host Zstd + device filters / block decoder,
    stenos.cpp:681-753) decode to the input; superblocks of 128 KiB .. 2 MiB."""
    ctx = api.Context()
    seen = {}
    for name, T, raw in dists.hybrid_cases(70000):
        for level in (2, 3, 6, 9):
            frame = np.frombuffer(ref.compress(raw, T, level=level), dtype=np.uint8)
            for c, k in dists.superblock_codes(frame, T, raw.size).items():
                seen[c] = seen.get(c, 0) + k
            assert ctx.decompress(frame, T, raw.size) == raw.tobytes(), (name, level)
    assert {2, 3, 4, 5} <= set(seen), seen
    # a damaged Zstd payload is invalid input, not a crash
    frame = np.frombuffer(ref.compress(dists.hybrid_cases(70000)[0][2], 4, level=3), dtype=np.uint8).copy()
    frame[12:16] ^= 0x5A  # the Zstd magic number of the first superblock's payload
    with pytest.raises(api.StenosError):
        ctx.decompress(frame, 4, 70000 * 4)


def _bucket_inputs(T, n_buckets, seed):
    """runs, noise, steps and constants side by side: buckets that code well, COPY buckets, all-same buckets"""
    rng = np.random.default_rng(seed)
    parts = []
    for b in range(n_buckets):
        k = b % 4
        if k == 0:
            v = dists.make("ramp_noise16", 256, T, seed=seed + b)
        elif k == 1:
            v = rng.integers(0, 256, 256 * T, dtype=np.uint8).view({2: np.int16, 4: np.int32, 8: np.int64}[T])
        elif k == 2:
            v = np.full(256, 42, dtype={2: np.int16, 4: np.int32, 8: np.int64}[T])
        else:
            v = dists.make("sparse_changes", 256, T, seed=seed + b)
        parts.append(raw_of(v))
    return np.concatenate(parts)


@pytest.mark.parametrize("T", [2, 4, 8])
def test_batched_bucket_encode_matches_the_per_bucket_call(T):
    """stenos_b200_compress_buckets_async == one stenos_private_compress_block per bucket with dst_size = the slot
    (cvector.hpp:1394-1420; room rule of SURVEY appendix C2), for full, partial-tail and selected (dirty) buckets."""
    ctx = api.Context()
    bb = 256 * T
    raw = _bucket_inputs(T, 13, seed=T)
    raw = np.concatenate([raw, raw[: bb // 2]])  # a last bucket of half the size
    n = (raw.size + bb - 1) // bb
    for stride, ids in ((bb + 16, None), (bb + 12, None), (bb + 16, np.array([3, 0, 13, 7, 7], dtype=np.uint32)), (bb // 2, None)):
        m = n if ids is None else ids.size
        slots = np.zeros(m * stride, dtype=np.uint8)
        sizes = np.zeros(m, dtype=np.uint32)
        res = np.zeros(2, dtype=np.uint64)
        ctx.compress_buckets_async(raw, T, bb, raw.size, ids, m, slots, stride, sizes, res)
        ctx.synchronize()
        overflow = False
        for i in range(m):
            b = int(i if ids is None else ids[i])
            bucket = raw[b * bb:(b + 1) * bb]
            want = run(port.compress_superblock, bucket, T, 1, stride)
            if isinstance(want, str):
                assert sizes[i] == 0, (T, stride, i)
                overflow = True
            else:
                assert sizes[i] == len(want) and slots[i * stride:i * stride + len(want)].tobytes() == want, (T, stride, i)
        assert (int(res[1]) & 1) == (1 if overflow else 0)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the compiled reference) writes the frames to compare with")
def test_forced_strategy_frames_equal_the_reference_where_it_picks_that_strategy():
    """stenos_b200_compress_strategy (device shuffle / shuffle + delta, host Zstd; stenos.cpp:617-656): wherever the
    reference codes EVERY superblock of a frame with strategy 3 or 4, the forced-strategy frame is the same bytes
    (superblock size and Zstd level of each level included); any forced frame decodes with the reference."""
    ctx = api.Context()
    ctx.set_threads(2)
    hits = {3: 0, 4: 0}
    for name, T, raw in dists.hybrid_cases(60000):
        for level in (3, 5, 6, 9):
            want = ref.compress(raw, T, level=level)
            codes = dists.superblock_codes(np.frombuffer(want, dtype=np.uint8), T, raw.size)
            for strat in (3, 4):
                got = ctx.compress_strategy(raw, T, level, strat)
                if set(codes) <= {strat}:
                    assert got == want, (name, level, strat)
                    hits[strat] += 1
                elif level == 5:
                    assert ref.decompress(got, T, raw.size) == raw.tobytes(), (name, level, strat)
                    assert ctx.decompress(np.frombuffer(got, dtype=np.uint8), T, raw.size) == raw.tobytes()
    assert hits[3] >= 1 and hits[4] >= 8, hits
    # strategy 5 (Zstd over the block stream, stenos.cpp:560-603) at level 2: the block encoder with the room of `bytes` per superblock
    hits5 = 0
    for name, T, raw in dists.hybrid_cases(60000):
        want = ref.compress(raw, T, level=2)
        codes = dists.superblock_codes(np.frombuffer(want, dtype=np.uint8), T, raw.size)
        got = ctx.compress_strategy(raw, T, 2, 5)
        if set(codes) <= {5, 1}:
            assert got == want, name
            hits5 += 1
        else:
            assert ref.decompress(got, T, raw.size) == raw.tobytes(), name
    assert hits5 >= 3, hits5
    assert run(ctx.compress_strategy, raw_of(np.arange(100000, dtype=np.int32)), 4, 3, 5) == "INVALID_PARAMETER"  # strategy 5: level 2 only
    assert run(ctx.compress_strategy, raw_of(np.arange(1000, dtype=np.int32)), 4, 1, 4) == "INVALID_PARAMETER"
    assert run(ctx.compress_strategy, raw_of(np.arange(1000, dtype=np.int32)), 4, 3, 6) == "INVALID_PARAMETER"
    assert run(ctx.compress_strategy, raw_of(np.arange(100000, dtype=np.int32)), 4, 3, 4, 100) == "DST_OVERFLOW"


@pytest.mark.parametrize("T", [3, 6])
def test_other_element_sizes_multi_superblock(T):
    """Element sizes 3 and 6 (first slice): superblocks of 130560 bytes (stenos.cpp:71-76), partial tail blocks, the
    < 128 byte Zstd tail, COPY superblocks -- frames identical to the oracle's, round trips, corrupt input rejected."""
    ctx = api.Context()
    sb = (131072 // (256 * T)) * 256 * T
    for name, n in (("ramp_noise16", sb // T + 300), ("random", sb // T + 5)):  # (the GPU fuzz test has the breadth; the emulator is slow)
        raw = raw_of(dists.make(name, n, T, seed=3))
        want = port.compress(raw, T)
        assert ctx.compress(raw, T) == want, (name, n)
        assert ctx.decompress(np.frombuffer(want, dtype=np.uint8), T, raw.size) == raw.tobytes(), (name, n)
        for ds in (len(want) - 1,):
            assert run(ctx.compress, raw, T, dst_size=ds) == run(port.compress, raw, T, dst_size=ds), (name, ds)
    bad = bytearray(port.compress(raw_of(dists.make("ramp_noise16", 3000, T, seed=4)), T))
    bad[8] = 9
    assert run(ctx.decompress, bytes(bad), T, 3000 * T) == "INVALID_INPUT"
    assert run(ctx.compress, raw_of(np.arange(495, dtype=np.uint8)), 9) == "INVALID_PARAMETER"


@pytest.mark.parametrize("T", [2, 4, 8])
def test_bucket_gather_fast_path_and_general_path(T):
    """stenos_b200_gather_decode_async on a cvector-style frame (one block per bucket): plane-coded buckets take the row
    decoder directly, COPY / LZ buckets and the partial last bucket the general superblock path; pairs mix both."""
    ctx = api.Context(block_shift=0)
    raw = _bucket_inputs(T, 21, seed=40 + T)
    raw = np.concatenate([raw, raw[: 256 * T // 2 + T]])  # a partial last bucket
    if T % 4 == 0:
        lz = raw_of(dists.make("lz_pairs", 256, T, seed=1))
        raw[5 * 256 * T:6 * 256 * T] = lz
    frame = np.frombuffer(port.compress(raw, T, block_shift=0), dtype=np.uint8)
    bb = 256 * T
    n_b = (raw.size + bb - 1) // bb
    offs = np.array(port.frame_index(frame, T), dtype=np.uint64)
    assert offs.size == n_b + 1
    ids = np.array(list(range(n_b)) + [3, 3, 0, n_b - 1, 7], dtype=np.uint32)
    out = np.zeros(ids.size * bb, dtype=np.uint8)
    res = np.zeros(2, dtype=np.uint64)
    ctx.gather_decode_async(frame, frame.size, T, bb, raw.size, offs, n_b, ids, ids.size, out, res)
    ctx.synchronize()
    assert res[1] == 0
    for k, b in enumerate(ids):
        want = raw[int(b) * bb:(int(b) + 1) * bb]
        assert out[k * bb:k * bb + want.size].tobytes() == want.tobytes(), (T, k, int(b))
    # a damaged bucket is reported, the others still decode
    bad = frame.copy()
    bad[int(offs[1]) + 4 + (T + 1) // 2] ^= 0xFF  # plane 0's first header byte of bucket 1
    ctx.gather_decode_async(bad, bad.size, T, bb, raw.size, offs, n_b, ids[:4], 4, out, res)
    ctx.synchronize()
    assert out[2 * bb:3 * bb].tobytes() == raw[2 * bb:3 * bb].tobytes()
