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
