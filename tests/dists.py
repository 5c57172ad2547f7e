"""Seeded input distributions shared by the oracle, golden-vector and GPU parity tests.

They are chosen to reach every branch of the block codec: all-same / raw planes, every bit width,
delta rows, RLE / delta-RLE rows, RLE-coded mins, LZ blocks (repeated values), LZ early exits,
COPY superblocks, partial tails.
"""
import numpy as np

_INT = {2: np.int16, 4: np.int32, 8: np.int64}


def _fit(a, T):
    """Reinterpret an int64 value array as T-byte little-endian elements (truncating)."""
    a = np.asarray(a, dtype=np.int64)
    if T in _INT:
        return a.astype(_INT[T])
    raw = a.astype("<i8").view(np.uint8).reshape(-1, 8)
    if T < 8:
        return np.ascontiguousarray(raw[:, :T]).reshape(-1)
    out = np.zeros((a.size, T), dtype=np.uint8)
    out[:, :8] = raw
    out[:, 8:] = raw[:, : T - 8] if T <= 16 else 0
    return out.reshape(-1)


def names():
    return [
        "zeros", "const", "sorted", "ramp_noise4", "ramp_noise16", "ramp_noise200", "random",
        "sparse_changes", "runs_random_len", "few_values", "repeat_period7", "smooth_sine",
        "steps", "lz_pairs", "lz_then_noise", "mostly_random_some_repeats", "byte_saw", "alternating",
        "big_jumps_rare", "low_entropy_bytes",
    ]


def make(name, n, T, seed=0):
    """Returns a numpy array holding n elements of T bytes (dtype intN for T in 2/4/8, else uint8 of n*T)."""
    rng = np.random.default_rng([names().index(name), seed, T])
    i = np.arange(n, dtype=np.int64)
    if name == "zeros":
        v = np.zeros(n, dtype=np.int64)
    elif name == "const":
        v = np.full(n, 0x0123456789ABCDEF if T == 8 else 0x1234567, dtype=np.int64)
    elif name == "sorted":
        v = i
    elif name == "ramp_noise4":
        v = 3 * i + rng.integers(0, 4, n)
    elif name == "ramp_noise16":
        v = 3 * i + rng.integers(0, 16, n) - 8
    elif name == "ramp_noise200":
        v = 7 * i + rng.integers(0, 200, n)
    elif name == "random":
        v = rng.integers(-(1 << 62), 1 << 62, n)
    elif name == "sparse_changes":
        v = np.cumsum((rng.random(n) < 0.03) * rng.integers(-1000, 1000, n))
    elif name == "runs_random_len":
        vals = rng.integers(-(1 << 40), 1 << 40, n // 20 + 2)
        lens = rng.integers(1, 60, vals.size)
        v = np.repeat(vals, lens)[:n]
        if v.size < n:
            v = np.concatenate([v, np.full(n - v.size, 5)])
    elif name == "few_values":
        v = rng.choice(np.array([0, 1, 255, 256, 65535, 1 << 20, -1, -300]), n)
    elif name == "repeat_period7":
        v = (rng.integers(-(1 << 50), 1 << 50, 7))[i % 7]
    elif name == "smooth_sine":
        v = np.rint(30000 * np.sin(i / 40.0) + 5 * rng.standard_normal(n)).astype(np.int64)
    elif name == "steps":
        v = (i // 37) * 1000003
    elif name == "lz_pairs":
        base = rng.integers(-(1 << 60), 1 << 60, n)
        v = base.copy()
        v[1::2] = base[0::2][: v[1::2].size]
    elif name == "lz_then_noise":
        base = rng.integers(-(1 << 60), 1 << 60, n)
        v = base.copy()
        k = int(rng.integers(1, max(2, min(200, n))))
        v[k:] = np.where(rng.random(n - k) < 0.5, v[:-k], v[k:])
    elif name == "mostly_random_some_repeats":
        v = rng.integers(-(1 << 60), 1 << 60, n)
        m = rng.random(n) < 0.15
        m[0] = False
        idx = np.nonzero(m)[0]
        v[idx] = v[idx - 1]
    elif name == "byte_saw":
        v = (i * 37) & 0xFF | (((i * 11) & 0xFF) << 8) | (((i // 3) & 0xFF) << 16) | (((i // 100) & 0xFF) << 24)
    elif name == "alternating":
        v = np.where(i & 1, 1 << 30, -(1 << 30)) + (i // 16)
    elif name == "big_jumps_rare":
        v = i + np.cumsum((rng.random(n) < 0.01) * (1 << 33))
    elif name == "low_entropy_bytes":
        b = rng.choice(np.array([0, 0, 0, 1, 2, 128, 255], dtype=np.uint8), n * T)
        return b.view(_INT[T]) if T in _INT else b
    else:
        raise ValueError(name)
    return _fit(v, T)


# ------------------------------------------------------------------------------------------------
# frames of level >= 2 (hybrid decoder): inputs on which the reference picks each of its Zstd strategies
# ------------------------------------------------------------------------------------------------
def hybrid_cases(n, seed=0):
    """(name, T, raw uint8 array): text -> code 2 (Zstd) and 3 (Zstd on the transposed input), steps -> 2 and 5 (blocks + Zstd),
    sensor / sine series -> 4 (transposed + delta + Zstd), small alphabet and walks -> 5."""
    from stenos_b200 import synth
    rng = np.random.default_rng(seed)
    text = np.frombuffer((b"the quick brown fox jumps over the lazy dog, " * (n * 4 // 45 + 2))[: n * 4], dtype=np.uint8)
    return [
        ("text", 4, text),
        ("steps", 4, (np.arange(n) // 1000 * 1000).astype(np.int32).view(np.uint8)),
        ("float64_sensor", 8, np.ascontiguousarray(synth.make("float64_sensor", n // 2)).view(np.uint8)),
        ("int16_sine", 2, np.ascontiguousarray(synth.make("int16_sine", n * 2 + 37)).view(np.uint8)),
        ("alphabet4", 4, rng.integers(0, 4, n * 4, dtype=np.uint8)),
        ("walk", 8, np.cumsum(rng.integers(-3, 4, n // 2)).astype(np.int64).view(np.uint8)),
    ]


def superblock_codes(frame, T, total):
    """Histogram of the superblock codes of a frame (host walk, stenos.cpp:1124-1143)."""
    first = 12 if frame[0] == 255 else 8
    if frame[0] == 255:
        sb = int.from_bytes(bytes(frame[8:12]), "little")
    else:
        bs = T * 256
        sb = (bs if bs > 131072 else (131072 // bs) * bs) << int(frame[0])
    at, out = first, {}
    for _ in range((total + sb - 1) // sb):
        out[int(frame[at])] = out.get(int(frame[at]), 0) + 1
        at += 4 + int.from_bytes(bytes(frame[at + 1:at + 4]), "little")
    return out
