"""Regenerates tests/golden/*.npz|json from the UNTOUCHED reference (oracle/_ref/libstenos_ref.so).

Run in the build container (needs `make -C oracle ref`):   python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these files --
outputs of the compiled reference on seeded inputs -- are what pins the byte stream.

  small.npz   : full input bytes + reference level-1 stream for small cases (every distribution,
                T in 2/4/8, sizes exercising full blocks, partial tails, < 128 byte Zstd/COPY tails)
  filters.npz : shuffle / shuffle+delta outputs of stenos::shuffle / stenos::delta
  large.json  : sha256 + length of the reference stream for multi-superblock seeded inputs
                (inputs are regenerated from tests/dists.py / stenos_b200/synth.py; their sha256 is
                stored too, so generator drift is detected instead of mis-reported as codec drift)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import dists  # noqa: E402
from oracle import ref  # noqa: E402
from stenos_b200 import synth  # noqa: E402


def raw_of(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def filter_input(T, n):
    """Seeded random-walk bytes (+ a ragged tail of < T bytes) for the filter fixtures."""
    rng = np.random.default_rng([17, T, n])
    a = (np.cumsum(rng.integers(-3, 4, n * T)).astype(np.int64) & 0xFF).astype(np.uint8)
    if n:
        a = np.concatenate([a, rng.integers(0, 256, int(rng.integers(0, T)), dtype=np.uint8)])
    return a


def main():
    small = {}
    for T in (2, 4, 8):
        for name in dists.names():
            for n in (256, 700, 33):
                raw = raw_of(dists.make(name, n, T, seed=7))
                key = "T%d_%s_%d" % (T, name, n)
                small[key + "_in"] = raw
                small[key + "_out"] = np.frombuffer(ref.compress(raw, T), dtype=np.uint8)
    # cvector-style buckets: one 256-element block with room T*256+16
    for T in (4, 8):
        for name in ("random", "lz_pairs", "sorted", "few_values"):
            raw = raw_of(dists.make(name, 256, T, seed=11))
            key = "bucket_T%d_%s" % (T, name)
            small[key + "_in"] = raw
            small[key + "_out"] = np.frombuffer(ref.compress_superblock(raw, T, room=T * 256 + 16), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "small.npz"), **small)

    filt = {}
    filt_large = []
    for T in (2, 4, 8):
        for n in (0, 1, 100, 2048 // T, 2048 // T + 1, 5000, 131072 // T, 262144 // T + 3):
            a = filter_input(T, n)
            sh = np.frombuffer(ref.shuffle(a, T), dtype=np.uint8)
            dl = np.frombuffer(ref.delta(sh), dtype=np.uint8)
            if a.size <= 48000:
                key = "T%d_%d" % (T, a.size)
                filt[key + "_in"] = a
                filt[key + "_shuffle"] = sh
                filt[key + "_shuffle_delta"] = dl
            else:
                filt_large.append(dict(kind="filter", T=T, n=n, in_sha256=sha(a), shuffle_sha256=sha(sh), shuffle_delta_sha256=sha(dl)))
    np.savez_compressed(os.path.join(HERE, "filters.npz"), **filt)

    large = list(filt_large)
    for T in (2, 4, 8):
        for name in dists.names():
            n = (3 * 131072 + 131072 // 2) // T + 77
            raw = raw_of(dists.make(name, n, T, seed=3))
            c = ref.compress(raw, T)
            large.append(dict(kind="dists", name=name, T=T, n=n, seed=3, in_sha256=sha(raw), out_len=len(c), out_sha256=sha(c)))
    for wl, n in (("int32_sorted", 1000000), ("int32_ramp_runs", 1 << 21), ("int64_ramp_runs", 1 << 20), ("int16_sine", 1 << 21),
                  ("float64_sensor", 1 << 19), ("float32_sensor", 1 << 20)):
        a = synth.make(wl, n)
        raw = raw_of(a)
        c = ref.compress(raw, a.itemsize)
        large.append(dict(kind="synth", name=wl, T=a.itemsize, n=n, seed=0, in_sha256=sha(raw), out_len=len(c), out_sha256=sha(c)))
    with open(os.path.join(HERE, "large.json"), "w") as f:
        json.dump(large, f, indent=1)
    print("small: %d arrays, filters: %d arrays, large: %d cases" % (len(small), len(filt), len(large)))


if __name__ == "__main__":
    main()
