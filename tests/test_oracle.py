"""Pins the CPU oracle (oracle/stenos_oracle.c) -- CPU only, no GPU.

1. against the committed golden vectors produced by the untouched reference (tests/golden/);
2. against the compiled reference itself (oracle/_ref) when it is present (differential fuzz),
   including the dst_size-dependent encoder decisions.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import dists
from oracle import port, ref
from stenos_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def raw_of(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def _small_cases():
    z = np.load(os.path.join(GOLD, "small.npz"))
    keys = sorted(k[:-3] for k in z.files if k.endswith("_in"))
    return z, keys


def test_golden_small_streams():
    z, keys = _small_cases()
    assert len(keys) > 150
    for k in keys:
        raw, want = z[k + "_in"], z[k + "_out"].tobytes()
        T = int(k.split("_")[1][1:]) if k.startswith("bucket") else int(k.split("_")[0][1:])
        if k.startswith("bucket"):
            got = port.compress_superblock(raw, T, room=T * 256 + 16)
        else:
            got = port.compress(raw, T)
            assert port.decompress(want, T, raw.size) == raw.tobytes(), k
        assert got == want, k


def test_golden_filters():
    z = np.load(os.path.join(GOLD, "filters.npz"))
    keys = sorted(k[:-3] for k in z.files if k.endswith("_in"))
    assert len(keys) >= 15
    for k in keys:
        T = int(k.split("_")[0][1:])
        a = z[k + "_in"]
        sh = port.shuffle(a, T)
        assert sh == z[k + "_shuffle"].tobytes(), k
        dl = port.delta(np.frombuffer(sh, dtype=np.uint8))
        assert dl == z[k + "_shuffle_delta"].tobytes(), k
        assert port.delta_inv(np.frombuffer(dl, dtype=np.uint8)) == sh, k
        assert port.unshuffle(np.frombuffer(sh, dtype=np.uint8), T) == a.tobytes(), k


def _large_cases():
    with open(os.path.join(GOLD, "large.json")) as f:
        return json.load(f)


def make_large_input(c):
    if c["kind"] == "dists":
        return raw_of(dists.make(c["name"], c["n"], c["T"], seed=c["seed"]))
    if c["kind"] == "synth":
        return raw_of(synth.make(c["name"], c["n"]))
    raise ValueError(c["kind"])


def test_golden_large_streams():
    n = 0
    for c in _large_cases():
        if c["kind"] == "filter":
            continue
        raw = make_large_input(c)
        assert sha(raw) == c["in_sha256"], ("input generator drifted", c["name"])
        got = port.compress(raw, c["T"])
        assert len(got) == c["out_len"], c
        assert sha(got) == c["out_sha256"], c
        assert port.decompress(got, c["T"], raw.size) == raw.tobytes()
        n += 1
    assert n >= 60


def test_readme_example_known_answer():
    # 1M sorted int32 at level 1 -> 70464 bytes, known first bytes
    a = np.arange(1000000, dtype=np.int32)
    c = port.compress(a, 4)
    assert len(c) == 70464
    assert c[:24].hex() == "0000093d000000000100090003008988888888888888fdff"


def test_exact_multiple_of_superblock_decodes():
    # deliberate divergence from the reference decoder bug
    a = synth.make("int32_ramp_runs", 2 * 32768)
    c = port.compress(a, 4)
    assert port.decompress(c, 4, a.nbytes) == a.tobytes()


def test_frame_index():
    a = synth.make("int32_ramp_runs", 5 * 32768 + 100)
    c = port.compress(a, 4)
    idx = port.frame_index(c, 4)
    assert len(idx) == 7 and idx[0] == 8 and idx[-1] == len(c)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_differential_vs_reference():
    bad = []
    for T in (2, 4, 8, 3, 12, 16):
        sb = port.lib().so_default_superblock(T)
        for name in dists.names():
            for n in (256, 256 * 7 + 13, 131072 // T + 100, 33):
                raw = raw_of(dists.make(name, n, T, seed=n + 1))
                c, c2 = ref.compress(raw, T), port.compress(raw, T)
                if c != c2:
                    bad.append((T, name, n))
                    continue
                if raw.size % sb:
                    assert ref.decompress(c2, T, raw.size) == raw.tobytes()
                assert port.decompress(c, T, raw.size) == raw.tobytes()
    assert not bad, bad


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_room_dependent_decisions_vs_reference():
    def run(fn, *a, **k):
        try:
            return fn(*a, **k)
        except RuntimeError as e:
            return str(e).split()[-1]

    for T in (2, 4, 8):
        for name in dists.names():
            raw = raw_of(dists.make(name, 256, T, seed=1))
            for room in (T * 256 + 16, T * 256 + 4, T * 256 + 64):
                assert run(ref.compress_superblock, raw, T, room=room) == run(port.compress_superblock, raw, T, room=room), (T, name, room)
            raw = raw_of(dists.make(name, 3000, T, seed=2))
            full = len(ref.compress(raw, T))
            for ds in (full + 40, full, full - 1, raw.size, 20, 7):
                assert run(ref.compress, raw, T, dst_size=ds) == run(port.compress, raw, T, dst_size=ds), (T, name, ds)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_filters_vs_reference():
    rng = np.random.default_rng(1)
    for T in (2, 3, 4, 8, 16):
        for n in (0, 5, 2047, 2048, 2049, 2051, 9999, 131072):
            a = rng.integers(0, 256, n, dtype=np.uint8)
            assert port.shuffle(a, T) == ref.shuffle(a, T)
            assert port.unshuffle(a, T) == ref.unshuffle(a, T)
            assert port.delta(a) == ref.delta(a)
            assert port.delta_inv(a) == ref.delta_inv(a)
