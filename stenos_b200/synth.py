"""Synthetic typed arrays for the BASELINE.json configs.

Every generator is closed-form in the GLOBAL element index (counter-based splitmix64 hash), so any
slice [start, start+count) can be produced independently by any rank / on any host and the CPU
reference and every GPU shard see identical data.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """Vectorised splitmix64 finaliser on a uint64 array."""
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _index(start, count):
    return np.arange(start, start + count, dtype=np.uint64)


def _hash(seed, idx):
    with np.errstate(over="ignore"):
        return splitmix64(idx ^ np.uint64(seed))


def _gauss(seed, idx):
    """Approximate N(0,1): sum of four 16-bit uniforms from one hash (Irwin-Hall), exact in float64."""
    h = _hash(seed, idx)
    s = np.zeros(idx.shape, dtype=np.float64)
    for k in range(4):
        s += ((h >> np.uint64(16 * k)) & np.uint64(0xFFFF)).astype(np.float64)
    return (s / 65536.0 - 2.0) * np.sqrt(3.0)


def _runs(v, start, count, value):
    """Elements [w*2^20+100000, w*2^20+400000) of every 2^20 window are set to `value`."""
    pos = (_index(start, count) & np.uint64((1 << 20) - 1)).astype(np.int64)
    v[(pos >= 100000) & (pos < 400000)] = value
    return v


def noisy_ramp_runs(dtype, start, count, seed=0):
    """Config 2 / config 4 (int64 half): v[i] = 3*i + (h(seed,i) mod 16) - 8, with runs of 42."""
    idx = _index(start, count)
    with np.errstate(over="ignore"):
        v = (idx * np.uint64(3) + (_hash(seed, idx) & np.uint64(15)) - np.uint64(8)).astype(np.uint64)
    v = v.astype(np.dtype(dtype).newbyteorder("=").str.replace("i", "u")).view(dtype)
    return _runs(v.copy(), start, count, 42)


def int16_sine(start, count, seed=0):
    """Config 3 / 4 (int16): 2000*sin(i/300) + 3*g."""
    idx = _index(start, count)
    x = 2000.0 * np.sin(idx.astype(np.float64) / 300.0) + 3.0 * _gauss(seed, idx)
    return np.rint(x).astype(np.int16)


def sensor_series(dtype, start, count, seed=0):
    """Config 3: x[i] = 20 + 5*sin(i/5000) + 1e-4*i + 0.01*g (float64 or float32)."""
    idx = _index(start, count)
    f = idx.astype(np.float64)
    x = 20.0 + 5.0 * np.sin(f / 5000.0) + 1e-4 * f + 0.01 * _gauss(seed, idx)
    return x.astype(dtype)


def sorted_int32(count):
    """Config 1 (README example): v[i] = i."""
    return np.arange(count, dtype=np.int32)


def make(name, count, start=0, seed=0):
    """Dispatch by workload name; returns a numpy array of `count` elements."""
    if name == "int32_ramp_runs":
        return noisy_ramp_runs(np.int32, start, count, seed)
    if name == "int64_ramp_runs":
        return noisy_ramp_runs(np.int64, start, count, seed)
    if name == "int16_sine":
        return int16_sine(start, count, seed)
    if name == "float64_sensor":
        return sensor_series(np.float64, start, count, seed)
    if name == "float32_sensor":
        return sensor_series(np.float32, start, count, seed)
    if name == "int32_sorted":
        return np.arange(start, start + count, dtype=np.int64).astype(np.int32)
    raise ValueError("unknown workload %r" % name)


# ------------------------------------------------------------------------------------------------
# The same closed forms evaluated ON THE DEVICE (torch), for shards too large to generate on the host
# in bench time (config 4: 32 GiB).  Integer generators are bit-identical to the numpy ones; the
# float-derived ones (sine, sensor) may differ in the last ulp of sin(), i.e. in a handful of
# elements -- parity checks therefore always read the device's actual input back.
# ------------------------------------------------------------------------------------------------
def _s64(c):
    c &= 0xFFFFFFFFFFFFFFFF
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr_t(x, k):
    return (x >> k) & ((1 << (64 - k)) - 1)


def _hash_t(seed, idx):
    z = (idx ^ _s64(seed)) + _s64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr_t(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr_t(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr_t(z, 31)


def make_torch(name, count, start=0, seed=0, device="cuda", chunk=1 << 26):
    """Device-side twin of make(): returns a torch tensor of `count` elements on `device`."""
    import torch

    dt = {"int32_ramp_runs": torch.int32, "int64_ramp_runs": torch.int64, "int16_sine": torch.int16, "float64_sensor": torch.float64,
          "float32_sensor": torch.float32}[name]
    out = torch.empty(count, dtype=dt, device=device)
    for c0 in range(0, count, chunk):
        n = min(chunk, count - c0)
        idx = torch.arange(start + c0, start + c0 + n, dtype=torch.int64, device=device)
        if name in ("int32_ramp_runs", "int64_ramp_runs"):
            v = idx * 3 + (_hash_t(seed, idx) & 15) - 8
            pos = idx & ((1 << 20) - 1)
            v = torch.where((pos >= 100000) & (pos < 400000), torch.full_like(v, 42), v)
            out[c0:c0 + n] = v.to(dt)  # int64 -> int32 keeps the low 32 bits (two's complement), like the numpy cast
        else:
            h = _hash_t(seed, idx)
            s = torch.zeros(n, dtype=torch.float64, device=device)
            for k in range(4):
                s += (_lsr_t(h, 16 * k) if k else h).bitwise_and(0xFFFF).to(torch.float64)
            g = (s / 65536.0 - 2.0) * (3.0 ** 0.5)
            f = idx.to(torch.float64)
            if name == "int16_sine":
                out[c0:c0 + n] = torch.round(2000.0 * torch.sin(f / 300.0) + 3.0 * g).to(dt)
            else:
                out[c0:c0 + n] = (20.0 + 5.0 * torch.sin(f / 5000.0) + 1e-4 * f + 0.01 * g).to(dt)
    return out
