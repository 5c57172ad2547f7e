#!/usr/bin/env python
"""The SURVEY.md section 8(d) configurations that the headline line of bench.py does not carry, measured on one GPU with
the inputs resident in HBM (CUDA events on the launching stream, inputs larger than L2):

  codec     level-1 compress / decompress for T = 2 (int16 sine) and T = 8 (int64 ramp + runs) at 1 GiB
  filters   shuffle / shuffle + delta / unshuffle / unshuffle + delta_inv on the 4 GiB float64 / float32 / int16 series
            (config 3), chunk = the superblock of level 3 (256 KiB)
  gather    cvector<int>-style frame (1 KiB buckets), 2^20 random buckets -> dense output (config 5)

`frac` = algorithmic bytes / time / MEASURED_PEAKS.json hbm_gbs.  Every row carries a `parity` flag: sampled superblocks /
chunks / buckets of exactly the data that was timed, compared with the CPU oracle (oracle/, test infrastructure -- the
checker, never the thing measured).  bench.py imports all_rows(); standalone:

    python tools/bench_rows.py [--mib 1024] [--rows codec,filters,gather] [--filter-mib 4096]
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

SB = 131072


def timeit(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def row_codec(T, name, nbytes, dev, pk, steps=5, warmup=3):
    from oracle import port

    d_src = synth.make_torch(name, nbytes // T, device=dev).view(torch.uint8)
    ctx = api.Context(level=1, stream=torch.cuda.current_stream())
    cap = api.bound(nbytes) + (1 << 20)
    d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    n_sb = (nbytes + SB - 1) // SB
    d_off = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    t_c = timeit(lambda: ctx.compress_async(d_src, T, nbytes, d_dst, cap, d_res, d_off), steps, warmup)
    c = int(d_res.cpu()[0])
    t_d = timeit(lambda: ctx.decompress_async(d_dst, T, c, d_out, nbytes, nbytes, d_res, d_off), steps, warmup)
    ok = int(d_res.cpu()[1]) == 0 and torch.equal(d_out, d_src)
    t_f = timeit(lambda: ctx.decompress_async(d_dst, T, c, d_out, nbytes, nbytes, d_res, None), steps, warmup)
    ok = ok and int(d_res.cpu()[1]) == 0 and torch.equal(d_out, d_src)
    offs = d_off.cpu().numpy()
    rng = np.random.RandomState(T)
    picks = sorted(set([0, n_sb - 1] + [int(x) for x in rng.randint(0, n_sb, size=16)]))
    for s in picks:
        raw = d_src[s * SB:(s + 1) * SB].cpu().numpy()
        got = d_dst[int(offs[s]):int(offs[s + 1])].cpu().numpy().tobytes()
        ok = ok and got == port.compress_superblock(raw, T, room=1 << 20)
    ctx.close()
    alg = nbytes + c
    return {"row": "codec", "T": T, "data": name, "bytes": nbytes, "algorithmic_bytes": alg, "ratio": nbytes / c,
            "compress_ms": t_c, "compress_GBps": nbytes / t_c / 1e6, "compress_frac": alg / t_c / 1e6 / pk,
            "decompress_ms": t_d, "decompress_GBps": nbytes / t_d / 1e6, "decompress_frac": alg / t_d / 1e6 / pk,
            "decompress_with_header_walk_ms": t_f, "decompress_with_header_walk_frac": alg / t_f / 1e6 / pk,
            "parity": bool(ok), "parity_superblocks": len(picks)}


def row_filters(T, name, nbytes, chunk, dev, pk, steps=5, warmup=3):
    from oracle import port

    d_src = synth.make_torch(name, nbytes // T, device=dev).view(torch.uint8)
    d_a = torch.empty_like(d_src)
    d_b = torch.empty_like(d_src)
    ctx = api.Context(level=1, stream=torch.cuda.current_stream())
    lib = ctx._lib
    out = {"row": "filters", "T": T, "data": name, "bytes": nbytes, "algorithmic_bytes": 2 * nbytes, "chunk": chunk}
    ok = True
    n_chunk = nbytes // chunk
    picks = [0, n_chunk // 3, n_chunk - 1]
    for wd in (0, 1):
        tag = "+delta" if wd else ""
        t_s = timeit(lambda: api.check(lib.stenos_b200_shuffle(ctx._h, T, nbytes, chunk, d_src.data_ptr(), d_a.data_ptr(), wd), "shuffle"), steps, warmup)
        t_u = timeit(lambda: api.check(lib.stenos_b200_unshuffle(ctx._h, T, nbytes, chunk, d_a.data_ptr(), d_b.data_ptr(), wd), "unshuffle"), steps, warmup)
        ok = ok and torch.equal(d_b, d_src)
        for k in picks:  # the reference's filter on the same chunk (stenos::shuffle, stenos::delta)
            raw = d_src[k * chunk:(k + 1) * chunk].cpu().numpy()
            want = port.shuffle(raw, T)
            if wd:
                want = port.delta(np.frombuffer(want, dtype=np.uint8))
            ok = ok and d_a[k * chunk:(k + 1) * chunk].cpu().numpy().tobytes() == want
        out["shuffle%s_ms" % tag] = t_s
        out["shuffle%s_frac" % tag] = 2 * nbytes / t_s / 1e6 / pk
        out["unshuffle%s_ms" % tag] = t_u
        out["unshuffle%s_frac" % tag] = 2 * nbytes / t_u / 1e6 / pk
    if T == 2:  # the plain byte delta and its inverse (no transpose; stenos::delta / delta_inv on every chunk), measured once
        t_d = timeit(lambda: api.check(lib.stenos_b200_delta(ctx._h, nbytes, chunk, d_src.data_ptr(), d_a.data_ptr()), "delta"), steps, warmup)
        t_i = timeit(lambda: api.check(lib.stenos_b200_delta_inv(ctx._h, nbytes, chunk, d_a.data_ptr(), d_b.data_ptr()), "delta_inv"), steps, warmup)
        ok = ok and torch.equal(d_b, d_src)
        for k in picks:
            raw = d_src[k * chunk:(k + 1) * chunk].cpu().numpy()
            ok = ok and d_a[k * chunk:(k + 1) * chunk].cpu().numpy().tobytes() == port.delta(raw)
        out["delta_ms"], out["delta_frac"] = t_d, 2 * nbytes / t_d / 1e6 / pk
        out["delta_inv_ms"], out["delta_inv_frac"] = t_i, 2 * nbytes / t_i / 1e6 / pk
    out["parity"] = bool(ok)
    ctx.close()
    return out


def row_gather(nbytes, n_ids, dev, pk, steps=5, warmup=3):
    T = 4
    # cvector<int>::serialize(): 12 byte header, one 1 KiB bucket per superblock; produced here by the device encoder with the same block size
    ctx = api.Context(level=1, stream=torch.cuda.current_stream(), block_shift=0)
    d_src = synth.make_torch("int32_ramp_runs", nbytes // T, device=dev).view(torch.uint8)
    cap = api.bound(nbytes) + nbytes // 256 + 64
    d_frame = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    n_b = (nbytes + 1023) // 1024
    d_off = torch.zeros(n_b + 1, dtype=torch.int64, device=dev)
    ctx.compress_async(d_src, T, nbytes, d_frame, cap, d_res, d_off)
    torch.cuda.synchronize()
    c = int(d_res.cpu()[0])
    ok = int(d_res.cpu()[1]) == 0
    g = torch.Generator(device="cpu").manual_seed(5)
    ids = torch.randint(0, n_b, (n_ids,), generator=g, dtype=torch.int32).to(dev)
    d_out = torch.empty(n_ids * 1024, dtype=torch.uint8, device=dev)
    t = timeit(lambda: ctx.gather_decode_async(d_frame, c, T, 1024, nbytes, d_off, n_b, ids, n_ids, d_out, d_res), steps, warmup)
    ok = ok and int(d_res.cpu()[1]) == 0
    want = d_src.view(-1, 1024)[ids.long()].reshape(-1)
    ok = ok and torch.equal(d_out, want)
    sizes = (d_off[1:] - d_off[:-1])[ids.long()].sum().item()
    ctx.close()
    alg = sizes + n_ids * 1024
    return {"row": "gather", "T": T, "buckets": n_b, "gathered": n_ids, "frame_bytes": c, "algorithmic_bytes": int(alg), "ms": t, "GBps_out": n_ids * 1024 / t / 1e6,
            "frac": alg / t / 1e6 / pk, "parity": bool(ok)}


def row_buckets(nbytes, dev, pk, steps=5, warmup=3):
    """cvector<int> write side: every 1 KiB bucket of the array compressed into its slot of 1 KiB + 16 in ONE launch
    (stenos_b200_compress_buckets_async), then the slots read back at random through the gather decoder."""
    T, bb = 4, 1024
    stride = bb + 16
    ctx = api.Context(level=1, stream=torch.cuda.current_stream())
    d_src = synth.make_torch("int32_ramp_runs", nbytes // T, device=dev).view(torch.uint8)
    n_b = nbytes // bb
    d_slots = torch.empty(n_b * stride, dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(n_b, dtype=torch.int32, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    t = timeit(lambda: ctx.compress_buckets_async(d_src, T, bb, nbytes, None, n_b, d_slots, stride, d_sizes, d_res), steps, warmup)
    ok = int(d_res.cpu()[1]) == 0
    sizes = d_sizes.cpu().numpy().astype(np.int64)
    slots = d_slots.cpu().numpy()
    src = d_src.cpu().numpy()
    from oracle import port  # the checker
    for i in np.random.default_rng(7).integers(0, n_b, 64):
        want = port.compress_superblock(src[i * bb:(i + 1) * bb], T, 1, stride)
        ok = ok and sizes[i] == len(want) and slots[i * stride:i * stride + len(want)].tobytes() == want
    # all buckets back through the gather decoder (offset of bucket i = i * stride)
    d_off = torch.arange(n_b + 1, dtype=torch.int64, device=dev) * stride
    ids = torch.arange(n_b, dtype=torch.int32, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ctx.gather_decode_async(d_slots, n_b * stride, T, bb, nbytes, d_off, n_b, ids, n_b, d_out, d_res)
    torch.cuda.synchronize()
    ok = ok and int(d_res.cpu()[1]) == 0 and torch.equal(d_out, d_src)
    ctx.close()
    alg = nbytes + int(sizes.sum())
    return {"row": "bucket_encode", "T": T, "buckets": int(n_b), "slot_stride": stride, "compressed_bytes": int(sizes.sum()), "algorithmic_bytes": int(alg), "ms": t,
            "GBps_in": nbytes / t / 1e6, "frac": alg / t / 1e6 / pk, "parity": bool(ok)}


def row_hybrid(name, T, nbytes, level=3, strategy=4):
    """Level >= 2, first slice of SURVEY 8 f1: forced-strategy compress (device shuffle + delta, host Zstd) and the hybrid decoder, host
    buffers in and out, next to the reference on the same host threads (its Zstd stage is the same library).  The frame must be the
    reference's own, byte for byte, where the reference picks that strategy for every superblock (config 3: it does)."""
    import time
    from oracle import ref  # comparison arm and checker

    def best(fn, n=3):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            r = fn()
            ts.append(time.perf_counter() - t0)
        return min(ts), r

    threads = os.cpu_count() or 1
    raw = np.ascontiguousarray(synth.make(name, nbytes // T)).view(np.uint8)
    ctx = api.Context()
    ctx.set_threads(threads)
    ctx.compress_strategy(raw[: 1 << 22], T, level, strategy)  # warm-up: context buffers, libzstd
    tc, frame = best(lambda: ctx.compress_strategy(raw, T, level, strategy))
    fr = np.frombuffer(frame, dtype=np.uint8)
    td, out = best(lambda: ctx.decompress(fr, T, raw.size))
    ok = out == raw.tobytes()
    trc, rframe = best(lambda: ref.compress(raw, T, level=level, threads=threads))
    raw2 = raw[: raw.size - T]  # (the reference cannot decode sizes that are a multiple of the superblock: SURVEY appendix C1)
    rf2 = ref.compress(raw2, T, level=level, threads=threads)
    trd, _ = best(lambda: ref.decompress(rf2, T, raw2.size, threads=threads))
    ctx.close()
    return {"row": "hybrid_level%d" % level, "data": name, "T": T, "bytes": int(raw.size), "strategy": strategy, "host_threads": threads,
            "frame_identical_to_reference": frame == rframe, "parity": bool(ok), "ratio": raw.size / len(frame), "compress_GBps": raw.size / tc / 1e9,
            "decompress_GBps": raw.size / td / 1e9, "reference_compress_GBps": raw.size / trc / 1e9, "reference_decompress_GBps": raw2.size / trd / 1e9,
            "note": "host Zstd bound on both sides; the device part is the filter (rows `filters`)"}


def row_generic(T, nbytes, dev, pk):
    """Element sizes 3, 5, 6, 7: the generic kernels through the synchronous reference-shaped calls on device pointers
    (one launch each; the call's own synchronisation is inside the time).  Parity: a 4 MiB prefix against the oracle's frame."""
    import time
    from oracle import port  # the checker
    n = nbytes // T
    i = torch.arange(n, dtype=torch.int64, device=dev)
    v = 3 * i + (i * 2654435761 % 16)
    raw = torch.stack([((v >> (8 * k)) & 255).to(torch.uint8) for k in range(T)], dim=1).reshape(-1).contiguous()
    ctx = api.Context()
    dst = torch.empty(api.bound(raw.numel()), dtype=torch.uint8, device=dev)
    out = torch.empty_like(raw)

    def best(fn, reps=3):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return min(ts), r

    ctx.compress_raw(raw, T, raw.numel(), dst, dst.numel())
    tc, r = best(lambda: ctx.compress_raw(raw, T, raw.numel(), dst, dst.numel()))
    ctx.decompress_raw(dst, T, r, out, out.numel())
    td, _ = best(lambda: ctx.decompress_raw(dst, T, r, out, out.numel()))
    ok = bool(torch.equal(out, raw))
    m = ((4 << 20) // (256 * T)) * 256 * T + 5 * T
    small = raw[:m].cpu().numpy()
    ok = ok and ctx.compress(small, T) == port.compress(small, T)
    ctx.close()
    alg = raw.numel() + int(r)
    return {"row": "codec_generic", "T": T, "bytes": int(raw.numel()), "ratio": raw.numel() / int(r), "algorithmic_bytes": alg, "compress_ms": tc * 1e3,
            "compress_GBps": raw.numel() / tc / 1e9, "compress_frac": alg / tc / 1e9 / pk, "decompress_ms": td * 1e3, "decompress_GBps": raw.numel() / td / 1e9,
            "decompress_frac": alg / td / 1e9 / pk, "parity": ok, "note": "encode_frame_kernel / decode_frame_kernel (generic in T), synchronous call on device pointers"}


def all_rows(dev, stream, pk, steps=5, warmup=3, codec_bytes=1 << 30, filter_bytes=4 << 30, gather_bytes=1 << 30):
    """Every row; a row that fails is reported as {"row": ..., "error": ...} instead of taking the bench line down with it."""
    jobs = []
    for T, name in ((2, "int16_sine"), (8, "int64_ramp_runs")):
        jobs.append(("codec", lambda T=T, name=name: row_codec(T, name, codec_bytes, dev, pk, steps, warmup)))
    for T, name in ((8, "float64_sensor"), (4, "float32_sensor"), (2, "int16_sine")):
        jobs.append(("filters", lambda T=T, name=name: row_filters(T, name, filter_bytes, 262144, dev, pk, steps, warmup)))
    jobs.append(("gather", lambda: row_gather(gather_bytes, 1 << 20, dev, pk, steps, warmup)))
    jobs.append(("bucket_encode", lambda: row_buckets(gather_bytes, dev, pk, steps, warmup)))
    jobs.append(("hybrid_level3", lambda: row_hybrid("float64_sensor", 8, 64 << 20)))
    for T in (3, 6):
        jobs.append(("codec_generic", lambda T=T: row_generic(T, 512 << 20, dev, pk)))
    rows = []
    with torch.cuda.stream(stream):
        for name, job in jobs:
            try:
                rows.append(job())
            except Exception as e:  # noqa: BLE001 -- reported, not hidden
                rows.append({"row": name, "error": "%s: %s" % (type(e).__name__, e)})
            torch.cuda.empty_cache()
    return rows


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--filter-mib", type=int, default=1024)
    ap.add_argument("--rows", default="codec,filters,gather,buckets,hybrid,generic")
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
                print(json.dumps(row_filters(T, name, args.filter_mib << 20, chunk, dev, pk)), flush=True)
    if "filters8" in rows:  # profiling: the T = 8 filters alone
        print(json.dumps(row_filters(8, "float64_sensor", args.filter_mib << 20, 262144, dev, pk)), flush=True)
    if "gather" in rows:
        print(json.dumps(row_gather(nbytes, 1 << 20, dev, pk)), flush=True)
    if "buckets" in rows:
        print(json.dumps(row_buckets(nbytes, dev, pk)), flush=True)
    if "generic" in rows:
        for T in (3, 5, 6, 7):
            print(json.dumps(row_generic(T, 512 << 20, dev, pk)), flush=True)
    if "hybrid" in rows:
        for name, T in (("float64_sensor", 8), ("int16_sine", 2)):
            print(json.dumps(row_hybrid(name, T, 256 << 20)), flush=True)


if __name__ == "__main__":
    main()
