#!/usr/bin/env python
"""Times the pieces of a decompress step separately (CUDA events): frame index, decode with a given index, both."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stenos_b200 import api, capi, synth
if os.environ.get("STENOS_B200_LIB"):
    capi.use_library(capi.load(os.environ["STENOS_B200_LIB"]))

def timeit(fn, n=5, w=2):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
    T = int(os.environ.get("TP_T", "4"))
    dev = torch.device("cuda:0")
    a = synth.make({2: "int16_sine", 4: "int32_ramp_runs", 8: "int64_ramp_runs"}[T], n * 4 // T)
    d_src = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    nbytes = d_src.numel()
    ctx = api.Context(level=1, stream=torch.cuda.current_stream())
    cap = api.bound(nbytes) + 16
    d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    n_sb = (nbytes + 131071) // 131072
    d_off = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
    d_off2 = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    t_c = timeit(lambda: ctx.compress_async(d_src, T, nbytes, d_dst, cap, d_res, d_off))
    c = int(d_res.cpu()[0])
    t_i = timeit(lambda: ctx.frame_index_async(d_dst, c, T, d_off2, n_sb + 1, d_res))
    acc = ctx.index_accepted()
    assert torch.equal(d_off, d_off2)
    t_d = timeit(lambda: ctx.decompress_async(d_dst, T, c, d_out, nbytes, nbytes, d_res, d_off))
    t_f = timeit(lambda: ctx.decompress_async(d_dst, T, c, d_out, nbytes, nbytes, d_res, None))
    if not os.environ.get("NOCHECK"):
        assert torch.equal(d_out, d_src)
    print("bytes %d csize %d | compress %.3f ms | index %.3f ms (accepted=%d) | decode(with index) %.3f ms | decode(full) %.3f ms" % (nbytes, c, t_c, t_i, acc, t_d, t_f))

main()
