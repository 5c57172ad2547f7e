#!/usr/bin/env python
"""End-to-end compress through the reference-facing call on pinned host buffers, for the staging modes of
STENOS_B200_ZERO_COPY (0: H2D copy + kernel + D2H copy; 1: kernel reads the pinned input in place; 2: writes the
pinned output in place; 3: both)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stenos_b200 import api, synth

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
    T = 4
    a = synth.make("int32_ramp_runs", n)
    host = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
    nbytes = host.numel()
    h_src = host.pin_memory()
    h_dst = torch.empty(api.bound(nbytes), dtype=torch.uint8).pin_memory()
    ref = None
    for mode in (0, 1, 2, 3):
        os.environ["STENOS_B200_ZERO_COPY"] = str(mode)
        ctx = api.Context(level=1)
        r = 0
        for _ in range(2):
            r = api.check(ctx.compress_raw(h_src, T, nbytes, h_dst, h_dst.numel()), "compress")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 5
        for _ in range(k):
            r = api.check(ctx.compress_raw(h_src, T, nbytes, h_dst, h_dst.numel()), "compress")
        dt = (time.perf_counter() - t0) / k
        out = bytes(h_dst[:r].numpy())
        if ref is None:
            ref = out
        print("zero-copy mode %d: %.2f ms  %.1f GB/s  csize %d  identical %s" % (mode, dt * 1e3, nbytes / dt / 1e9, r, out == ref), flush=True)
        ctx.close()

main()
