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
