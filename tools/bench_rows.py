#!/usr/bin/env python
"""Measures the SURVEY.md section 8 rows that bench.py's headline line does not carry, on one GPU, with the
inputs resident in HBM (CUDA events on the launching stream, inputs larger than L2):

  codec     level-1 compress / decompress for T = 2 (int16 sine), 4 (int32 ramp + runs), 8 (int64 ramp + runs)
  filters   shuffle / shuffle + delta / unshuffle / unshuffle + delta for T = 2, 4, 8 (config 3 data), chunk = superblock
            sizes of levels 2 / 3 / 5 (128 / 256 / 512 KiB)
  gather    cvector<int>-style frame (1 KiB buckets), 2^20 random buckets -> dense output (config 5)

One JSON object per line on stdout; `frac` = algorithmic bytes / time / MEASURED_PEAKS.json hbm_gbs.

    python tools/bench_rows.py [--mib 1024] [--rows codec,filters,gather]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stenos_b200 import api, capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("STENOS_B200_LIB"):  # experiment builds (tools/build_variant.py)
    capi.use_library(capi.load(os.environ["STENOS_B200_LIB"]))


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, n=5, w=3):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def to_dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)


def row_codec(T, name, nbytes, dev, pk):
    a = synth.make(name, nbytes // T)
    d_src = to_dev(a, dev)
    ctx = api.Context(level=1, stream=torch.cuda.current_stream())
    cap = api.bound(nbytes) + 16
    d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    n_sb = (nbytes + 131071) // 131072
    d_off = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    t_c = timeit(lambda: ctx.compress_async(d_src, T, nbytes, d_dst, cap, d_res, d_off))
    c = int(d_res.cpu()[0])
    t_d = timeit(lambda: ctx.decompress_async(d_dst, T, c, d_out, nbytes, nbytes, d_res, d_off))
    t_f = timeit(lambda: ctx.decompress_async(d_dst, T, c, d_out, nbytes, nbytes, d_res, None))
    assert int(d_res.cpu()[1]) == 0 and torch.equal(d_out, d_src)
    ctx.close()
    alg = nbytes + c
    return {"row": "codec", "T": T, "data": name, "bytes": nbytes, "ratio": nbytes / c,
            "compress_ms": t_c, "compress_GBps": nbytes / t_c / 1e6, "compress_frac": alg / t_c / 1e6 / pk,
            "decompress_ms": t_d, "decompress_GBps": nbytes / t_d / 1e6, "decompress_frac": alg / t_d / 1e6 / pk,
            "decompress_with_frame_walk_ms": t_f}


def row_filters(T, name, nbytes, chunk, dev, pk):
    a = synth.make(name, nbytes // T)
    d_src = to_dev(a, dev)
    d_a = torch.empty_like(d_src)
    d_b = torch.empty_like(d_src)
    ctx = api.Context(level=1, stream=torch.cuda.current_stream())
    lib = ctx._lib  # the library the context was made by (STENOS_B200_LIB experiments included)
    out = {"row": "filters", "T": T, "data": name, "bytes": nbytes, "chunk": chunk}
    for wd in (0, 1):
        tag = "+delta" if wd else ""
        t_s = timeit(lambda: api.check(lib.stenos_b200_shuffle(ctx._h, T, nbytes, chunk, d_src.data_ptr(), d_a.data_ptr(), wd), "shuffle"))
        t_u = timeit(lambda: api.check(lib.stenos_b200_unshuffle(ctx._h, T, nbytes, chunk, d_a.data_ptr(), d_b.data_ptr(), wd), "unshuffle"))
        assert torch.equal(d_b, d_src)
        out["shuffle%s_ms" % tag] = t_s
        out["shuffle%s_frac" % tag] = 2 * nbytes / t_s / 1e6 / pk
        out["unshuffle%s_ms" % tag] = t_u
        out["unshuffle%s_frac" % tag] = 2 * nbytes / t_u / 1e6 / pk
    ctx.close()
    return out


def row_gather(nbytes, n_ids, dev, pk):
    T = 4
    a = synth.make("int32_ramp_runs", nbytes // T)
    # cvector<int>::serialize(): 12 byte header, one 1 KiB bucket per superblock; produced here by the device encoder with the same block size
    ctx = api.Context(level=1, stream=torch.cuda.current_stream(), block_shift=0)
    d_src = to_dev(a, dev)
    cap = api.bound(nbytes) + nbytes // 256 + 64
    d_frame = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    n_b = (nbytes + 1023) // 1024
    d_off = torch.zeros(n_b + 1, dtype=torch.int64, device=dev)
    ctx.compress_async(d_src, T, nbytes, d_frame, cap, d_res, d_off)
    torch.cuda.synchronize()
    c = int(d_res.cpu()[0])
    assert int(d_res.cpu()[1]) == 0
    g = torch.Generator(device="cpu").manual_seed(5)
    ids = torch.randint(0, n_b, (n_ids,), generator=g, dtype=torch.int32).to(dev)
    d_out = torch.empty(n_ids * 1024, dtype=torch.uint8, device=dev)
    t = timeit(lambda: ctx.gather_decode_async(d_frame, c, T, 1024, nbytes, d_off, n_b, ids, n_ids, d_out, d_res))
    assert int(d_res.cpu()[1]) == 0
    want = d_src.view(-1, 1024)[ids.long()].reshape(-1)
    assert torch.equal(d_out, want)
    sizes = (d_off[1:] - d_off[:-1])[ids.long()].sum().item()
    ctx.close()
    alg = sizes + n_ids * 1024
    return {"row": "gather", "T": T, "buckets": n_b, "gathered": n_ids, "frame_bytes": c, "ms": t, "GBps_out": n_ids * 1024 / t / 1e6, "frac": alg / t / 1e6 / pk}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--rows", default="codec,filters,gather")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    pk = peak()
    nbytes = args.mib << 20
    rows = args.rows.split(",")
    if "codec" in rows:
        for T, name in ((2, "int16_sine"), (4, "int32_ramp_runs"), (8, "int64_ramp_runs")):
            print(json.dumps(row_codec(T, name, nbytes, dev, pk)), flush=True)
    if "filters" in rows:
        for T, name in ((2, "int16_sine"), (4, "float32_sensor"), (8, "float64_sensor")):
            for chunk in (131072, 262144, 524288):
                print(json.dumps(row_filters(T, name, nbytes, chunk, dev, pk)), flush=True)
    if "gather" in rows:
        print(json.dumps(row_gather(nbytes, 1 << 20, dev, pk)), flush=True)


if __name__ == "__main__":
    main()
