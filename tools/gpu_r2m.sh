#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b13_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b13_pytest.log; tail -4 gpurun_out/b13_pytest.log
python - <<'P'
# throughput of the generic element sizes (1 GiB-ish, host buffers are too slow to matter: device-side timing through the sync API on device pointers)
import numpy as np, torch, time, sys
sys.path.insert(0, '.')
from stenos_b200 import api
for T in (3, 6):
    n = (512 << 20) // T
    i = torch.arange(n, dtype=torch.int64, device='cuda')
    v = (3 * i + (i * 2654435761 % 16)).to(torch.int64)
    raw = torch.stack([((v >> (8 * k)) & 255).to(torch.uint8) for k in range(T)], dim=1).reshape(-1).contiguous()
    ctx = api.Context()
    dst = torch.empty(api.bound(raw.numel()), dtype=torch.uint8, device='cuda')
    out = torch.empty_like(raw)
    r = ctx.compress_raw(raw, T, raw.numel(), dst, dst.numel()); torch.cuda.synchronize()
    t0 = time.perf_counter(); r = ctx.compress_raw(raw, T, raw.numel(), dst, dst.numel()); torch.cuda.synchronize(); tc = time.perf_counter() - t0
    d = ctx.decompress_raw(dst, T, r, out, out.numel()); torch.cuda.synchronize()
    t0 = time.perf_counter(); d = ctx.decompress_raw(dst, T, r, out, out.numel()); torch.cuda.synchronize(); td = time.perf_counter() - t0
    print("T=%d: %d MiB, ratio %.2f, compress %.1f GB/s, decompress %.1f GB/s, roundtrip %s" % (T, raw.numel() >> 20, raw.numel() / r, raw.numel() / tc / 1e9, raw.numel() / td / 1e9, bool(torch.equal(out, raw))))
P
