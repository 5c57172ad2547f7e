#!/usr/bin/env python
"""bench.py -- level-1 compress / decompress throughput of the B200 path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch: compress the
rank's shard of the synthetic int32 array (noisy ramp + runs, BASELINE.json configs[1]: 1 GiB per GPU),
then (separately timed) decompress it.  `value` is whole-job compress GB/s of UNCOMPRESSED bytes with
the input resident in HBM; `e2e` is the same metric through the reference-facing C ABI call
(stenos_compress_generic) on pinned HOST buffers, copies inside the timed region.

N > 1 (torchrun, one process per GPU): weak scaling -- every rank owns 1 GiB of one N GiB frame,
partitioned on superblock boundaries; only the segment byte lengths are all-gathered (NCCL).

--impl reference times the UNTOUCHED reference (oracle/_ref/libstenos_ref.so, AVX2, all host threads)
on a bounded sample of the same workload, on rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "level-1 compress GB/s of uncompressed bytes, whole job (decompress GB/s, ratio and %HBM roofline in extra keys)"
WORKLOAD = "int32_ramp_runs"
T = 4
SHARD_ELEMS = 1 << 28  # 1 GiB of int32 per GPU


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(sample_elems, threads, steps, warmup):
    """The reference's own CPU implementation on `threads` host threads: GB/s compress / decompress."""
    import ctypes as C
    from oracle import ref
    from stenos_b200 import synth

    L = ref.lib()
    a = synth.make(WORKLOAD, sample_elems)
    src = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    n = src.size
    cap = L.stenos_bound(n)
    dst = np.empty(cap, dtype=np.uint8)
    ctx = L.stenos_make_context()
    L.stenos_set_level(ctx, 1)
    L.stenos_set_threads(ctx, threads)
    ctimes, r = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = L.stenos_compress_generic(ctx, src.ctypes.data, T, n, dst.ctypes.data, cap)
        dt = time.perf_counter() - t0
        if i >= warmup:
            ctimes.append(dt)
    assert not ref.has_error(r)
    # decode: the reference decoder rejects exact multiples of the superblock size:
    # time it on one 256-element block fewer
    n2 = n - T * 256
    r2 = L.stenos_compress_generic(ctx, src.ctypes.data, T, n2, dst.ctypes.data, cap)
    out = np.empty(n2, dtype=np.uint8)
    dtimes = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        d = L.stenos_decompress_generic(ctx, dst.ctypes.data, T, r2, out.ctypes.data, n2)
        dt = time.perf_counter() - t0
        if i >= warmup:
            dtimes.append(dt)
    assert d == n2 and out.tobytes() == src[:n2].tobytes()
    L.stenos_destroy_context(ctx)
    return {"compress_GBps": n / np.mean(ctimes) / 1e9, "decompress_GBps": n2 / np.mean(dtimes) / 1e9, "ratio": n / r, "bytes": n,
            "ms_per_step": float(np.mean(ctimes) * 1e3)}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 1 << 26  # 256 MiB of int32 per step
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libstenos_ref.so was not built in the container"}))
        return
    m = cpu_reference_arm(sample, cores, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": m["compress_GBps"], "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "%s level 1, %d MiB sample of the 1 GiB/GPU array, CPU threads=%d" % (WORKLOAD, m["bytes"] >> 20, cores)},
        "decompress_GBps": m["decompress_GBps"], "ratio": m["ratio"],
        "cpu_baseline": {"value": m["compress_GBps"], "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": "%d MiB of %s, stenos_set_threads=%d, AVX2 build of /root/reference" % (m["bytes"] >> 20, WORKLOAD, cores)},
        "e2e": {"value": m["compress_GBps"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--shard-elems", type=int, default=SHARD_ELEMS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly ONE line, the JSON: whatever libraries print there meanwhile (NCCL's version banner at the
    # first collective) is sent to stderr, and the real stdout comes back for the last print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from stenos_b200 import api, build, distributed, synth

    build.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.shard_elems
    nbytes = n * T
    frame_bytes = nbytes * world
    a = synth.make(WORKLOAD, n, start=rank * n)
    host = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
    stream = torch.cuda.current_stream()
    ctx = api.Context(level=1, stream=stream)
    seg = distributed.SegmentCodec(ctx, T, frame_bytes)
    d_src = host.to(dev)
    # device-resident destination: the worst case of the frame plus 1 MiB, so that every superblock -- the last one too --
    # provably has the room in which the reference's dst-room checks are inert and the whole
    # frame is one launch of the stream encoder; the e2e leg below uses stenos_bound(bytes) exactly
    cap = seg.capacity(nbytes) + (1 << 20)
    d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(2, dtype=torch.int64, device=dev)
    n_sb = (nbytes + seg.sb - 1) // seg.sb
    d_off = torch.zeros(n_sb + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def compress_step():
        if world == 1:
            ctx.compress_async(d_src, T, nbytes, d_dst, cap, d_res, d_off)
        else:
            seg.compress_async(d_src, nbytes, d_dst, cap, d_res, d_off)
            # only the segment byte length leaves the GPU: 8 bytes per rank, all-gathered
            sizes = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(sizes, d_res[:1])

    csize_holder = {}

    def decompress_step():
        if world == 1:
            ctx.decompress_async(d_dst, T, csize_holder["c"], d_out, nbytes, nbytes, d_res, None)  # includes the on-device header walk
        else:
            seg.decompress_async(d_dst, csize_holder["c"], nbytes, d_off, d_out, d_res)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        l0 = api.kernel_launches()
        t0 = time.perf_counter()
        for e0, e1 in evs:
            e0.record(stream)
            fn()
            e1.record(stream)
        sync_all()
        wall = time.perf_counter() - t0
        per = [e0.elapsed_time(e1) for e0, e1 in evs]  # ms, device time on the launching stream
        total_ms = evs[0][0].elapsed_time(evs[-1][1])
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), per, api.kernel_launches() - l0, wall

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    c_ms, c_per, c_launch, _ = timed(compress_step, args.steps, args.warmup)
    res = d_res.cpu().numpy()
    assert res[1] == 0, "device error bits %d" % res[1]
    csize = int(res[0])
    csize_holder["c"] = csize
    d_ms, d_per, d_launch, _ = timed(decompress_step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    assert d_res.cpu().numpy()[1] == 0
    assert torch.equal(d_out, d_src), "round trip mismatch"

    # ---- e2e: the reference-facing call on pinned host buffers (H2D + kernels + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        h_src = host.pin_memory()
        h_dst = torch.empty(api.bound(nbytes), dtype=torch.uint8).pin_memory()
        ectx = api.Context(level=1, stream=stream)
        r = 0
        for _ in range(2):
            r = api.check(ectx.compress_raw(h_src, T, nbytes, h_dst, h_dst.numel()), "stenos_compress_generic")
        sync_all()
        t0 = time.perf_counter()
        esteps = max(3, min(args.steps, 5))
        for _ in range(esteps):
            r = api.check(ectx.compress_raw(h_src, T, nbytes, h_dst, h_dst.numel()), "stenos_compress_generic")
        sync_all()
        et = torch.tensor([(time.perf_counter() - t0) / esteps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        e2e = {"value": nbytes * world / float(et.item()) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(r) + 16,
               "call": "stenos_compress_generic(ctx, pinned host src, 4, bytes, pinned host dst, stenos_bound(bytes))"}
        ectx.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    c_step_ms = c_ms / args.steps
    d_step_ms = d_ms / args.steps
    k_ms = float(np.mean(c_per))  # device time of one compress step on the launching stream (2 memsets + the encode kernel)
    alg_bytes = nbytes + csize
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("encode_stream_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": frame_bytes / (c_step_ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": c_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "%d GiB int32 noisy ramp + runs per GPU (BASELINE.json configs[1]), level 1, one frame of %d GiB" % (nbytes >> 30, frame_bytes >> 30)
                   if nbytes >= (1 << 30) else "%d MiB int32 noisy ramp + runs per GPU, level 1" % (nbytes >> 20),
                   "bytesoftype": T, "superblock": seg.sb, "l2": "inputs larger than L2 (no flush needed)", "device_dst_capacity": "stenos_bound(bytes) + 1 MiB",
                   "parallelism": "superblock ranges per GPU, segment sizes all-gathered" if world > 1 else "1 GPU"},
        "decompress_GBps": frame_bytes / (d_step_ms * 1e-3) / 1e9, "decompress_ms_per_step": d_step_ms,
        "ratio": nbytes / csize, "compressed_bytes_per_gpu": csize,
        "roofline": {"bound": "hbm", "kernel": "encode_stream_kernel<4,640>", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms},
        "roofline_decompress": {"bound": "hbm", "kernel": "decode_pairs_kernel<4> (+ index_scan/merge/fill: the step walks the frame headers on the device)", "achieved": alg_bytes / (float(np.mean(d_per)) * 1e-3) / 1e9, "peak": peak,
                                "unit": "GB/s", "frac": alg_bytes / (float(np.mean(d_per)) * 1e-3) / 1e9 / peak, "traffic": None},
        "gpu_launches": int(c_launch + d_launch), "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        from oracle import ref
        if ref.available():
            cores = os.cpu_count() or 1
            m = cpu_reference_arm(1 << 26, cores, 3, 1)
            m1 = cpu_reference_arm(1 << 24, 1, 3, 1)
            line["cpu_baseline"] = {"value": m["compress_GBps"], "unit": "GB/s", "cores": cores, "kind": "reference",
                                    "sample": "256 MiB of the same array; stenos_set_threads=%d" % cores,
                                    "decompress_GBps": m["decompress_GBps"], "single_thread_compress_GBps": m1["compress_GBps"],
                                    "single_thread_decompress_GBps": m1["decompress_GBps"], "ratio": m["ratio"]}
        else:
            from oracle import port
            s = np.ascontiguousarray(synth.make(WORKLOAD, 1 << 22)).view(np.uint8)
            t0 = time.perf_counter()
            port.compress(s, T)
            line["cpu_baseline"] = {"value": s.size / (time.perf_counter() - t0) / 1e9, "unit": "GB/s", "cores": 1, "kind": "port", "sample": "16 MiB, scalar C port"}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
