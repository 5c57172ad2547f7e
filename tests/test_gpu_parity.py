"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI of
libstenos_b200.so, against the CPU oracle on the same seeded inputs, against the committed golden
vectors of the reference, and -- at BASELINE.json sizes -- through size independent properties.
Bar: bit exact (integer / byte work)."""
import hashlib
import time
import json
import os

import numpy as np
import pytest

import dists
from oracle import port
from stenos_b200 import api, capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", autouse=True)
def native_library():
    # the nvcc build, never the emulator: fail loudly if it is missing
    capi.use_library(capi.load())
    assert capi.lib().stenos_b200_build_target() == b"sm_100a"
    yield


def raw_of(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def run(fn, *a, **k):
    try:
        return fn(*a, **k)
    except api.StenosError as e:
        return e.name
    except RuntimeError as e:
        return {"-6": "DST_OVERFLOW", "-2": "SRC_OVERFLOW", "-4": "INVALID_INPUT"}.get(str(e).split()[-1], str(e))


def test_readme_example_known_answer():
    a = np.arange(1000000, dtype=np.int32)
    c = api.compress(a, 4)
    assert len(c) == 70464
    assert c[:24].hex() == "0000093d000000000100090003008988888888888888fdff"
    assert api.decompress(c, 4, a.nbytes) == a.tobytes()


def test_golden_small_streams():
    z = np.load(os.path.join(GOLD, "small.npz"))
    keys = sorted(k[:-3] for k in z.files if k.endswith("_in"))
    ctx = api.Context()
    for k in keys:
        raw, want = z[k + "_in"], z[k + "_out"].tobytes()
        if k.startswith("bucket"):
            T = int(k.split("_")[1][1:])
            assert ctx.compress_block(raw, T, room=T * 256 + 16) == want, k
            assert ctx.decompress_block(want, T, raw.size) == raw.tobytes(), k
        else:
            T = int(k.split("_")[0][1:])
            assert ctx.compress(raw, T) == want, k
            assert ctx.decompress(want, T, raw.size) == raw.tobytes(), k


def test_golden_large_streams():
    with open(os.path.join(GOLD, "large.json")) as f:
        cases = json.load(f)
    ctx = api.Context()
    n = 0
    for c in cases:
        if c["kind"] == "filter":
            continue
        raw = raw_of(dists.make(c["name"], c["n"], c["T"], seed=c["seed"]) if c["kind"] == "dists" else synth.make(c["name"], c["n"]))
        assert sha(raw) == c["in_sha256"]
        got = ctx.compress(raw, c["T"])
        assert len(got) == c["out_len"] and sha(got) == c["out_sha256"], c
        assert ctx.decompress(got, c["T"], raw.size) == raw.tobytes()
        n += 1
    assert n >= 60


def test_golden_filters():
    z = np.load(os.path.join(GOLD, "filters.npz"))
    ctx = api.Context()
    for k in sorted(k[:-3] for k in z.files if k.endswith("_in")):
        T = int(k.split("_")[0][1:])
        a = z[k + "_in"]
        assert ctx.shuffle(a, T) == z[k + "_shuffle"].tobytes(), k
        assert ctx.shuffle(a, T, 0, True) == z[k + "_shuffle_delta"].tobytes(), k
        assert ctx.unshuffle(z[k + "_shuffle"], T) == a.tobytes(), k
        assert ctx.unshuffle(z[k + "_shuffle_delta"], T, 0, True) == a.tobytes(), k


@pytest.mark.parametrize("T", [2, 4, 8, 3, 5, 6, 7])  # 3, 5, 6, 7: the generic kernels
def test_fuzz_vs_oracle(T):
    ctx = api.Context()
    for name in dists.names():
        for n in (256, 256 * 7 + 13, 131072 // T + 100, 33, (5 * 131072) // T + 999):
            raw = raw_of(dists.make(name, n, T, seed=n + 1))
            want = port.compress(raw, T)
            assert ctx.compress(raw, T) == want, (name, n)
            assert ctx.decompress(want, T, raw.size) == raw.tobytes(), (name, n)


def test_edge_cases():
    ctx = api.Context()
    assert ctx.compress(b"", 4) == port.compress(np.zeros(0, np.uint8), 4)
    assert ctx.decompress(ctx.compress(b"", 4), 4, 0) == b""
    for n in (1, 31, 32, 33, 63):  # tail superblock < 128 bytes -> Zstd or COPY
        raw = raw_of(dists.make("ramp_noise4", n, 4, seed=n))
        want = port.compress(raw, 4)
        assert ctx.compress(raw, 4) == want
        assert ctx.decompress(want, 4, raw.size) == raw.tobytes()
    raw = raw_of(dists.make("sorted", 2 * 32768, 4))  # exact multiple of the superblock (reference decoder bug C1)
    c = ctx.compress(raw, 4)
    assert c == port.compress(raw, 4) and ctx.decompress(c, 4, raw.size) == raw.tobytes()
    raw = raw_of(dists.make("random", 40000, 8, seed=1))  # incompressible -> COPY superblocks
    c = ctx.compress(raw, 8)
    assert c == port.compress(raw, 8) and ctx.decompress(c, 8, raw.size) == raw.tobytes()
    assert api.compress(raw, 8, level=0) == port.compress(raw, 8, level=0)
    assert run(api.compress, raw, 8, level=3) == "INVALID_PARAMETER"
    assert api.compress(raw[:39999], 3) == port.compress(raw[:39999], 3)  # element size 3: the generic kernels
    assert run(api.compress, raw[:39996], 9) == "INVALID_PARAMETER"  # not built: no CPU fallback


def test_room_dependent_decisions():
    ctx = api.Context()
    for T in (2, 4, 8, 3, 5, 6, 7):
        for name in dists.names():
            raw = raw_of(dists.make(name, 256, T, seed=1))
            for room in (T * 256 + 16, T * 256 + 4, T * 256 + 64):
                assert run(ctx.compress_block, raw, T, room=room) == run(port.compress_superblock, raw, T, room=room), (T, name, room)
            raw = raw_of(dists.make(name, 3000, T, seed=2))
            full = len(port.compress(raw, T))
            for ds in (full + 40, full, full - 1, raw.size, 20, 7):
                assert run(ctx.compress, raw, T, dst_size=ds) == run(port.compress, raw, T, dst_size=ds), (T, name, ds)


def test_corrupt_streams_are_rejected():
    ctx = api.Context()
    raw = raw_of(dists.make("ramp_noise16", 50000, 4, seed=4))
    c = bytearray(port.compress(raw, 4))
    assert run(ctx.decompress, bytes(c[: len(c) // 2]), 4, raw.size) in ("SRC_OVERFLOW", "INVALID_INPUT")
    assert run(ctx.decompress, bytes(c), 4, raw.size - 1) == "DST_OVERFLOW"
    rng = np.random.default_rng(0)
    for _ in range(50):
        c3 = bytearray(c)
        for k in rng.integers(8, len(c), 6):
            c3[k] ^= int(rng.integers(1, 256))
        assert isinstance(run(ctx.decompress, bytes(c3), 4, raw.size), (bytes, str))
    assert ctx.decompress(bytes(c), 4, raw.size) == raw.tobytes()  # the context is still healthy


def test_filters_vs_oracle_with_chunks():
    ctx = api.Context()
    for wl, T in (("float64_sensor", 8), ("float32_sensor", 4), ("int16_sine", 2)):
        a = raw_of(synth.make(wl, (1 << 20) // T + 3))
        for chunk in (131072, 262144, 524288, 2080 * T // 2, 16 * T * 1021):  # the last two: odd multiples of 16 elements (T=2: mid-group stream starts)
            pieces = [a[i:i + chunk] for i in range(0, a.size, chunk)]
            sh = b"".join(port.shuffle(p, T) for p in pieces)
            shd = b"".join(port.delta(np.frombuffer(port.shuffle(p, T), dtype=np.uint8)) for p in pieces)
            assert ctx.shuffle(a, T, chunk) == sh
            assert ctx.shuffle(a, T, chunk, True) == shd
            assert ctx.unshuffle(np.frombuffer(sh, dtype=np.uint8), T, chunk) == a.tobytes()
            assert ctx.unshuffle(np.frombuffer(shd, dtype=np.uint8), T, chunk, True) == a.tobytes()
            assert ctx.delta(a, chunk) == b"".join(port.delta(p) for p in pieces)


def test_device_resident_async_api_and_index():
    import torch

    dev = torch.device("cuda:0")
    a = synth.make("int32_ramp_runs", 1 << 22)
    want = port.compress(a, 4)
    ctx = api.Context(stream=torch.cuda.current_stream())
    d_src = torch.from_numpy(raw_of(a).copy()).to(dev)
    cap = api.bound(a.nbytes)
    d_dst = torch.zeros(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    n_sb = (a.nbytes + 131071) // 131072
    d_off = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
    ctx.compress_async(d_src, 4, a.nbytes, d_dst, cap, d_res, d_off)
    torch.cuda.synchronize()
    res = d_res.cpu().numpy()
    assert res[1] == 0 and res[0] == len(want)
    assert d_dst[: len(want)].cpu().numpy().tobytes() == want
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), port.frame_index(want, 4))
    # decode with the index produced by the encoder, and with the on-device header walk
    for offs in (d_off, None):
        d_out = torch.zeros(a.nbytes, dtype=torch.uint8, device=dev)
        ctx.decompress_async(d_dst, 4, len(want), d_out, a.nbytes, a.nbytes, d_res, offs)
        torch.cuda.synchronize()
        assert d_res.cpu().numpy()[1] == 0
        assert d_out.cpu().numpy().tobytes() == a.tobytes()
    # frame index of a foreign (oracle produced) frame
    d_f = torch.from_numpy(np.frombuffer(want, dtype=np.uint8).copy()).to(dev)
    d_off2 = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
    assert ctx.frame_index_async(d_f, len(want), 4, d_off2, n_sb + 1, d_res) == n_sb
    torch.cuda.synchronize()
    assert np.array_equal(d_off2.cpu().numpy().astype(np.uint64), port.frame_index(want, 4))


def test_cvector_bucket_gather():
    import torch

    dev = torch.device("cuda:0")
    n_buckets = 4096
    a = synth.make("int32_ramp_runs", n_buckets * 256)
    frame = port.compress(a, 4, block_shift=0)  # what cvector<int>::serialize() produces (12 byte header, 1 KiB buckets)
    ctx = api.Context(stream=torch.cuda.current_stream(), block_shift=0)
    assert ctx.compress(a, 4) == frame
    d_f = torch.from_numpy(np.frombuffer(frame, dtype=np.uint8).copy()).to(dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    d_off = torch.zeros(n_buckets + 1, dtype=torch.int64, device=dev)
    assert ctx.frame_index_async(d_f, len(frame), 4, d_off, n_buckets + 1, d_res) == n_buckets
    rng = np.random.default_rng(1)
    ids = rng.integers(0, n_buckets, 10000).astype(np.uint32)
    d_ids = torch.from_numpy(ids.view(np.int32).copy()).to(dev)
    d_out = torch.zeros(ids.size * 1024, dtype=torch.uint8, device=dev)
    ctx.gather_decode_async(d_f, len(frame), 4, 1024, a.nbytes, d_off, n_buckets, d_ids, ids.size, d_out, d_res)
    torch.cuda.synchronize()
    assert d_res.cpu().numpy()[1] == 0
    got = d_out.cpu().numpy().view(np.int32).reshape(-1, 256)
    assert np.array_equal(got, a.reshape(-1, 256)[ids])


def test_full_size_roundtrip_properties():
    """BASELINE.json config 2 at full size (1 GiB int32): device resident round trip; the stream is
    checked against the oracle on sampled superblocks (superblocks are independent) and through a
    checksum of the decoded output."""
    import torch

    dev = torch.device("cuda:0")
    n = 1 << 28
    a = synth.make("int32_ramp_runs", n)
    ctx = api.Context(stream=torch.cuda.current_stream())
    d_src = torch.from_numpy(raw_of(a)).to(dev)
    cap = api.bound(a.nbytes)
    d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    n_sb = a.nbytes // 131072
    d_off = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
    ctx.compress_async(d_src, 4, a.nbytes, d_dst, cap, d_res, d_off)
    torch.cuda.synchronize()
    total, err = (int(x) for x in d_res.cpu().numpy())
    assert err == 0
    offs = d_off.cpu().numpy()
    assert offs[0] == 8 and offs[-1] == total and np.all(np.diff(offs) > 4)
    head = d_dst[:8].cpu().numpy().tobytes()
    assert head == bytes([0]) + int(a.nbytes).to_bytes(7, "little")
    rng = np.random.default_rng(7)
    for s in [0, 1, n_sb - 1] + list(rng.integers(0, n_sb, 24)):
        s = int(s)
        got = d_dst[int(offs[s]): int(offs[s + 1])].cpu().numpy().tobytes()
        want = port.compress_superblock(a[s * 32768: (s + 1) * 32768], 4, room=1 << 20)
        assert got == want, s
    d_out = torch.empty(a.nbytes, dtype=torch.uint8, device=dev)
    ctx.decompress_async(d_dst, 4, total, d_out, a.nbytes, a.nbytes, d_res, d_off)
    torch.cuda.synchronize()
    assert d_res.cpu().numpy()[1] == 0
    assert torch.equal(d_out, d_src)


def test_parallel_frame_index_on_adversarial_frames():
    """The device frame index re-synchronises on the header chain in parallel; payloads that imitate
    headers must not fool it (merge-time verification, serial fallback)."""
    import torch

    dev = torch.device("cuda:0")
    ctx = api.Context(stream=torch.cuda.current_stream(), block_shift=0)
    n_sb = 40000
    rng = np.random.default_rng(1)
    a = rng.integers(-2**31, 2**31, 256 * n_sb).astype(np.int32)
    a[: 256 * 15000: 3] = 7
    frames = [port.compress(a, 4, block_shift=0, dst_size=a.nbytes + 8 * n_sb + 64)]
    for fake in (1, 0x00001001, 0x00000406):
        b = np.full(256 * n_sb, fake, dtype=np.int32)
        frames.append(port.compress(b, 4, level=0, block_shift=0, dst_size=b.nbytes + 8 * n_sb + 64))
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    for i, frame in enumerate(frames):
        d_f = torch.from_numpy(np.frombuffer(frame, dtype=np.uint8).copy()).to(dev)
        d_off = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
        assert ctx.frame_index_async(d_f, len(frame), 4, d_off, n_sb + 1, d_res) == n_sb
        torch.cuda.synchronize()
        assert d_res.cpu().numpy()[1] == 0
        assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), port.frame_index(frame, 4))
        if i == 0:
            assert ctx.index_accepted() == 1  # an ordinary frame is indexed by the parallel path, no serial fallback


def _mixed(T, n_sb, seed, tail_elems):
    per = 131072 // T
    kinds = ["random", "ramp_noise16", "random", "lz_then_noise", "random", "random", "sparse_changes", "mostly_random_some_repeats", "const"]
    parts = [raw_of(dists.make(kinds[(i + seed) % len(kinds)], per, T, seed=seed + i)) for i in range(n_sb)]
    if tail_elems:
        parts.append(raw_of(dists.make("random" if seed & 1 else "ramp_noise16", tail_elems, T, seed=seed)))
    return np.concatenate(parts)


@pytest.mark.parametrize("T,n_sb,tail", [(4, 700, 255), (8, 450, 37), (2, 380, 0), (4, 33, 100)])
def test_stream_encoder_mixed_superblocks(T, n_sb, tail):
    """encode_stream_kernel (ample dst room -> every superblock through the barrier-free pipeline): long runs of
    incompressible superblocks (ring wrap, waits for flushes, COPY), LZ blocks, constant data, a partial tail --
    bit exact against the oracle, and the same bytes as the barrier kernel (STENOS_B200_LEGACY_ENCODER)."""
    raw = _mixed(T, n_sb, seed=T + n_sb, tail_elems=tail)
    room = raw.size + 300000
    want = port.compress(raw, T, dst_size=room)
    ctx = api.Context()
    os.environ["STENOS_B200_PIPELINE_CHUNK"] = "0"  # one launch for the whole frame
    try:
        l0 = api.kernel_launches()
        got = ctx.compress(raw, T, dst_size=room)
        assert api.kernel_launches() - l0 == 1
    finally:
        del os.environ["STENOS_B200_PIPELINE_CHUNK"]
    assert got == want
    assert ctx.decompress(got, T, raw.size) == raw.tobytes()
    # host buffers, pipelined: chunks of whole superblocks as independent segments, copies overlapped with the kernels
    os.environ["STENOS_B200_PIPELINE_CHUNK"] = str(8 << 20)
    try:
        l0 = api.kernel_launches()
        assert ctx.compress(raw, T, dst_size=room) == want
        n_launch = api.kernel_launches() - l0
        assert n_launch >= 2 if raw.size >= (16 << 20) else n_launch == 1
    finally:
        del os.environ["STENOS_B200_PIPELINE_CHUNK"]
    os.environ["STENOS_B200_LEGACY_ENCODER"] = "1"
    try:
        assert api.Context().compress(raw, T, dst_size=room) == want
    finally:
        del os.environ["STENOS_B200_LEGACY_ENCODER"]


def test_stream_encoder_many_tiny_superblocks():
    ctx = api.Context(block_shift=0)
    raw = np.concatenate([raw_of(dists.make(n, 256 * 4000, 4, seed=i)) for i, n in enumerate(("sparse_changes", "random", "lz_pairs", "const", "ramp_noise16"))]
                         + [raw_of(dists.make("sorted", 77, 4))])
    room = raw.size * 2 + 4096
    want = port.compress(raw, 4, block_shift=0, dst_size=room)
    assert ctx.compress(raw, 4, dst_size=room) == want
    assert api.decompress(want, 4, raw.size) == raw.tobytes()


def test_unchanged_cvector_header_is_a_drop_in(tmp_path):
    """the reference's unchanged stenos/cvector.hpp compiled against libstenos_b200.so (oracle/Makefile,
    target cvector; tests/cpp/cvector_dropin.cpp) behaves like the same program linked against the reference: same
    element values after push_back / random writes / iteration, and the same serialized bytes (the cvector room rule
    T*256+16 per bucket)."""
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "cvector_dropin_ref")
    b200_bin = os.path.join(ROOT, "oracle", "_ref", "cvector_dropin_b200")
    if not (os.path.exists(ref_bin) and os.path.exists(b200_bin)):
        pytest.skip("drop-in binaries not built (make -C oracle cvector needs /root/reference)")
    a, b = str(tmp_path / "ref.bin"), str(tmp_path / "b200.bin")
    ra = subprocess.run([ref_bin, a], capture_output=True, text=True, timeout=300)
    rb = subprocess.run([b200_bin, b], capture_output=True, text=True, timeout=600)
    assert ra.returncode == 0, ra.stdout + ra.stderr
    assert rb.returncode == 0, rb.stdout + rb.stderr
    assert ra.stdout == rb.stdout
    assert open(a, "rb").read() == open(b, "rb").read()


# ------------------------------------------------------------------------------------------------
# round 2
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,name", [(2, "int16_sine"), (8, "int64_ramp_runs")])
def test_full_size_sampled_superblocks_other_element_sizes(T, name):
    """BASELINE.json configs at 1 GiB for the element sizes the headline does not cover: device resident round trip,
    the frame header, the superblock index, and sampled superblocks byte for byte against the oracle."""
    import torch

    dev = torch.device("cuda:0")
    nbytes = 1 << 30
    d_src = synth.make_torch(name, nbytes // T, device=dev).view(torch.uint8)
    ctx = api.Context(stream=torch.cuda.current_stream())
    cap = api.bound(nbytes) + (1 << 20)
    d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    n_sb = nbytes // 131072
    d_off = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
    ctx.compress_async(d_src, T, nbytes, d_dst, cap, d_res, d_off)
    torch.cuda.synchronize()
    total, err = (int(x) for x in d_res.cpu().numpy())
    assert err == 0
    offs = d_off.cpu().numpy()
    assert offs[0] == 8 and offs[-1] == total and np.all(np.diff(offs) > 4)
    assert d_dst[:8].cpu().numpy().tobytes() == bytes([0]) + int(nbytes).to_bytes(7, "little")
    rng = np.random.default_rng(T)
    for s in [0, 1, n_sb - 1] + list(rng.integers(0, n_sb, 29)):
        s = int(s)
        raw = d_src[s * 131072:(s + 1) * 131072].cpu().numpy()
        got = d_dst[int(offs[s]): int(offs[s + 1])].cpu().numpy().tobytes()
        assert got == port.compress_superblock(raw, T, room=1 << 20), s
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for index in (d_off, None):  # with the encoder's index, and walking the headers on the device
        d_out.zero_()
        ctx.decompress_async(d_dst, T, total, d_out, nbytes, nbytes, d_res, index)
        torch.cuda.synchronize()
        assert d_res.cpu().numpy()[1] == 0
        assert torch.equal(d_out, d_src)


def test_decoder_accepts_raw_block_marker_252():
    """Blocks stored raw behind marker 252 only appear in time limited streams of the reference (block_compress.h:2118-2123),
    which this encoder never writes; the decoder must read them.  Hand-built streams, checked against the oracle's decoder."""
    rng = np.random.default_rng(252)
    for T in (2, 4, 8):
        blk = T * 256
        blocks = [raw_of(dists.make(n, 256, T, seed=i)) for i, n in enumerate(("ramp_noise16", "random", "const", "lz_pairs", "sorted"))]
        payload = b""
        for i, b in enumerate(blocks):
            if i % 2 == 1:
                payload += bytes([252]) + b.tobytes()
            else:
                sb = port.compress_superblock(b, T, room=1 << 16)  # [1][csize:3][the block's stream]
                assert sb[0] == 1
                payload += sb[4:]
        tail = raw_of(dists.make("ramp_noise4", 40, T, seed=9))
        sbt = port.compress_superblock(np.concatenate([blocks[0], tail]), T, room=1 << 16)
        assert sbt[0] == 1
        first_len = len(port.compress_superblock(blocks[0], T, room=1 << 16)) - 4
        payload_with_tail = payload + sbt[4 + first_len:]  # the partial tail block of a two-block superblock
        for body, nraw in ((payload, b"".join(b.tobytes() for b in blocks)), (payload_with_tail, b"".join(b.tobytes() for b in blocks) + tail.tobytes())):
            frame = bytes([0]) + len(nraw).to_bytes(7, "little") + bytes([1]) + len(body).to_bytes(3, "little") + body
            assert port.decompress(frame, T, len(nraw)) == nraw
            assert api.decompress(frame, T, len(nraw)) == nraw
        bad = bytearray(bytes([0]) + (5 * blk).to_bytes(7, "little") + bytes([1]) + len(payload).to_bytes(3, "little") + payload)
        del bad[-7:]  # truncated inside the last block
        bad[9:12] = (len(payload) - 7).to_bytes(3, "little")
        assert run(api.decompress, bytes(bad), T, 5 * blk) in ("INVALID_INPUT", "SRC_OVERFLOW")


def test_flow_encoder_spill_path_on_the_gpu(monkeypatch):
    """Staging rings at their legal minimum: with poorly compressible data every warp keeps moving pieces to the HBM spill
    slots (sb_flow.cuh).  The streams must be the oracle's."""
    monkeypatch.setenv("STENOS_B200_FLOW_RING", "1")
    monkeypatch.setenv("STENOS_B200_PIPELINE_CHUNK", "0")
    for T in (2, 4, 8):
        raw = _mixed(T, 600, seed=7 * T, tail_elems=77)
        room = raw.size + 400000
        want = port.compress(raw, T, dst_size=room)
        ctx = api.Context()
        assert ctx.compress(raw, T, dst_size=room) == want
        assert ctx.decompress(want, T, raw.size) == raw.tobytes()


def test_encoder_next_to_a_kernel_that_holds_most_sms():
    """Co-residency: the persistent encoder must not assume that all of its CTAs run at once.  A kernel on another stream
    holds 140 SMs' shared memory for a while; the encoder is launched behind it with the full grid and must finish with
    the oracle's stream (round 1 handed every CTA its first superblock statically and could wait forever here)."""
    import torch

    dev = torch.device("cuda:0")
    T = 4
    raw = _mixed(T, 1500, seed=5, tail_elems=0)
    room = raw.size + 400000
    want = port.compress(raw, T, dst_size=room)
    d_src = torch.from_numpy(raw).to(dev)
    d_dst = torch.empty(room, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream()
    ctx = api.Context(stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    api.check(capi.lib().stenos_b200_test_occupy(side.cuda_stream, 140, 200, 300_000_000), "occupy")  # 0.3 s
    time.sleep(0.02)
    t0 = time.perf_counter()
    ctx.compress_async(d_src, T, raw.size, d_dst, room, d_res, None)
    torch.cuda.current_stream().synchronize()
    dt = time.perf_counter() - t0
    side.synchronize()
    total, err = (int(x) for x in d_res.cpu().numpy())
    assert err == 0 and d_dst[:total].cpu().numpy().tobytes() == want
    assert dt < 5.0


def test_pipelined_host_decompress_and_cached_contexts():
    """stenos_decompress_generic on large host buffers runs in chunks (H2D, decode, D2H overlapped); stenos_compress /
    stenos_decompress reuse one context per thread."""
    T = 4
    a = synth.make("int32_ramp_runs", (192 << 20) // 4 + 1000)
    raw = raw_of(a)
    ctx = api.Context()
    c = ctx.compress(raw, T)
    assert hashlib.sha256(c).hexdigest() == hashlib.sha256(port.compress(raw, T)).hexdigest()
    l0 = api.kernel_launches()
    assert ctx.decompress(c, T, raw.size) == raw.tobytes()
    assert api.kernel_launches() - l0 >= 4  # one decode launch per 32 MiB chunk
    bad = bytearray(c)
    bad[len(bad) // 2] ^= 0x55
    bad[len(bad) // 2 + 1] ^= 0x55
    r = run(ctx.decompress, bytes(bad), T, raw.size)
    assert r in ("INVALID_INPUT", "SRC_OVERFLOW") or isinstance(r, bytes)  # never a crash; a payload flip may decode to other data
    small = raw[: 300000]
    for _ in range(3):
        cs = api.compress(small, T)
        assert cs == port.compress(small, T)
        assert api.decompress(cs, T, small.size) == small.tobytes()


def _ref_program(name):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = os.path.join(root, "oracle", "_ref", name)
    if not os.path.exists(p):
        pytest.skip("%s not built (make -C oracle reftests needs /root/reference)" % name)
    return p


def test_reference_own_comp_decomp_test_against_the_boundary():
    """The reference's tests/tests_comp_decomp.cpp (its test_vector(): sentinels behind dst and behind the decoded buffer,
    'an error implies dst_size < stenos_bound', round trip, shrinking dst_size) compiled against libstenos_b200.so; the
    sweep is cut to T in {2, 4, 8} and levels 0..1 (tests/cpp/ref_comp_decomp_cut.cpp)."""
    import subprocess

    r = subprocess.run([_ref_program("ref_comp_decomp_cut_b200")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "ref_comp_decomp_cut ok" in r.stdout


@pytest.mark.skipif(os.environ.get("STENOS_B200_LONG_TESTS") != "1", reason="minutes of per-bucket calls: set STENOS_B200_LONG_TESTS=1")
def test_reference_own_cvector_test_against_the_boundary():
    """The reference's tests/test_cvector.cpp, unchanged (cvector<size_t>, <int>, move-only and atomic payloads, 16 reader
    threads), compiled against libstenos_b200.so."""
    import subprocess

    r = subprocess.run([_ref_program("ref_test_cvector_main_b200")], capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "ref_test_cvector ok" in r.stdout


def test_hybrid_decoder_reads_the_reference_frames_of_every_level():
    """Frames written by the REFERENCE at levels 2..9 (superblock codes 2..5, superblocks of 128 KiB .. 2 MiB) decode to
    the input: Zstd on the host (1 and 4 threads), inverse filters and the block decoder on the GPU; host and device
    resident frames / outputs.  (stenos.cpp:681-753)"""
    import torch
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref (the compiled reference) writes the frames")
    ctx = api.Context()
    seen = {}
    for name, T, raw in dists.hybrid_cases(1 << 20):
        for level in range(2, 10):
            frame = np.frombuffer(ref.compress(raw, T, level=level), dtype=np.uint8)
            for c, k in dists.superblock_codes(frame, T, raw.size).items():
                seen[c] = seen.get(c, 0) + k
            ctx.set_threads(4 if level % 2 else 1)
            assert ctx.decompress(frame, T, raw.size) == raw.tobytes(), (name, level)
            if level in (3, 6):
                d_frame = torch.from_numpy(frame.copy()).cuda()
                d_out = torch.zeros(raw.size, dtype=torch.uint8, device="cuda")
                assert ctx.decompress_raw(d_frame, T, frame.size, d_out, raw.size) == raw.size
                assert bytes(d_out.cpu().numpy()) == raw.tobytes(), (name, level, "device")
    assert {1, 2, 3, 4, 5} <= set(seen), seen
    frame = np.frombuffer(ref.compress(dists.hybrid_cases(1 << 18)[2][2], 8, level=5), dtype=np.uint8).copy()
    frame[12:16] ^= 0x5A  # the Zstd magic number of the first superblock's payload
    with pytest.raises(api.StenosError):
        ctx.decompress(frame, 8, (1 << 18) * 4)


@pytest.mark.parametrize("T", [2, 4, 8])
def test_batched_bucket_encode_matches_the_per_bucket_call(T):
    """stenos_b200_compress_buckets_async (device resident, one launch) == stenos_private_compress_block per bucket with
    dst_size = the slot (cvector.hpp:1394-1420, room rule of SURVEY appendix C2); its slots are then read back through
    stenos_b200_gather_decode_async with the slot offsets as the index."""
    import torch

    dev = torch.device("cuda:0")
    bb = 256 * T
    n = 3000
    a = synth.make({2: "int16_sine", 4: "int32_ramp_runs", 8: "int64_ramp_runs"}[T], n * 256)
    raw = raw_of(a).copy()
    rng = np.random.default_rng(T)
    for b in rng.integers(0, n, 300):  # incompressible buckets -> COPY
        raw[b * bb:(b + 1) * bb] = rng.integers(0, 256, bb, dtype=np.uint8)
    ctx = api.Context(stream=torch.cuda.current_stream())
    d_src = torch.from_numpy(raw).to(dev)
    stride = bb + 16
    d_slots = torch.zeros(n * stride, dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(n, dtype=torch.int32, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    l0 = api.kernel_launches()
    ctx.compress_buckets_async(d_src, T, bb, raw.size, None, n, d_slots, stride, d_sizes, d_res)
    torch.cuda.synchronize()
    assert api.kernel_launches() - l0 == 1
    assert d_res.cpu().numpy()[1] == 0
    slots, sizes = d_slots.cpu().numpy(), d_sizes.cpu().numpy()
    host = api.Context()
    for i in list(range(0, n, 7)) + [n - 1]:
        bucket = raw[i * bb:(i + 1) * bb]
        want = port.compress_superblock(bucket, T, 1, stride)
        assert sizes[i] == len(want) and slots[i * stride:i * stride + len(want)].tobytes() == want, (T, i)
        if i % 49 == 0:
            assert host.compress_block(bucket, T, super_block_size=bb, room=stride) == want
    # dirty buckets only, then random access to the slots
    ids = rng.permutation(n)[:500].astype(np.uint32)
    d_ids = torch.from_numpy(ids.view(np.int32).copy()).to(dev)
    d_slots2 = torch.zeros(ids.size * stride, dtype=torch.uint8, device=dev)
    d_sizes2 = torch.zeros(ids.size, dtype=torch.int32, device=dev)
    ctx.compress_buckets_async(d_src, T, bb, raw.size, d_ids, ids.size, d_slots2, stride, d_sizes2, d_res)
    torch.cuda.synchronize()
    s2 = d_slots2.cpu().numpy()
    for k in range(0, ids.size, 11):
        i = int(ids[k])
        assert s2[k * stride:k * stride + sizes[i]].tobytes() == slots[i * stride:i * stride + sizes[i]].tobytes()
    d_off = (torch.arange(n + 1, dtype=torch.int64, device=dev) * stride).contiguous()
    d_out = torch.zeros(ids.size * bb, dtype=torch.uint8, device=dev)
    ctx.gather_decode_async(d_slots, n * stride, T, bb, raw.size, d_off, n, d_ids, ids.size, d_out, d_res)
    torch.cuda.synchronize()
    assert d_res.cpu().numpy()[1] == 0
    assert np.array_equal(d_out.cpu().numpy().reshape(-1, bb), raw.reshape(-1, bb)[ids])


def test_forced_strategy_frames_equal_the_reference_where_it_picks_that_strategy():
    """stenos_b200_compress_strategy on the GPU (shuffle / shuffle + delta kernels, host Zstd; stenos.cpp:617-656): the
    config-3 series (float64 / float32 sensor, int16 sine -- the reference codes them as strategy 4 at every level >= 3)
    give the reference's frames byte for byte; device resident input and output too."""
    import torch
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref (the compiled reference) writes the frames to compare with")
    ctx = api.Context()
    ctx.set_threads(8)
    hits = 0
    for name, T in (("float64_sensor", 8), ("float32_sensor", 4), ("int16_sine", 2)):
        raw = raw_of(synth.make(name, (3 << 20) // T + 256))
        for level in (3, 4, 5, 7, 9):
            want = ref.compress(raw, T, level=level, threads=4)
            codes = dists.superblock_codes(np.frombuffer(want, dtype=np.uint8), T, raw.size)
            got = ctx.compress_strategy(raw, T, level, 4)
            if set(codes) <= {4}:
                assert got == want, (name, level)
                hits += 1
            assert ref.decompress(got, T, raw.size) == raw.tobytes()
            assert ctx.decompress(np.frombuffer(got, dtype=np.uint8), T, raw.size) == raw.tobytes()
        d_src = torch.from_numpy(raw.copy()).cuda()
        d_dst = torch.zeros(api.bound(raw.size), dtype=torch.uint8, device="cuda")
        r = capi.lib().stenos_b200_compress_strategy(ctx._h, d_src.data_ptr(), T, raw.size, d_dst.data_ptr(), d_dst.numel(), 5, 4)
        assert r == len(ctx.compress_strategy(raw, T, 5, 4)) and bytes(d_dst[:r].cpu().numpy()) == ctx.compress_strategy(raw, T, 5, 4)
    assert hits >= 12, hits
    # strategy 5 at level 2 (Zstd over the block stream, stenos.cpp:560-603): the config-2 / config-4 integer arrays
    hits5 = 0
    for name, T in (("int32_ramp_runs", 4), ("int64_ramp_runs", 8), ("int16_sine", 2)):
        raw = raw_of(synth.make(name, (5 << 20) // T + 999))
        want = ref.compress(raw, T, level=2, threads=4)
        codes = dists.superblock_codes(np.frombuffer(want, dtype=np.uint8), T, raw.size)
        got = ctx.compress_strategy(raw, T, 2, 5)
        if set(codes) <= {5, 1}:
            assert got == want, name
            hits5 += 1
        assert ref.decompress(got, T, raw.size) == raw.tobytes()
        assert ctx.decompress(np.frombuffer(got, dtype=np.uint8), T, raw.size) == raw.tobytes()
    assert hits5 >= 2, hits5
