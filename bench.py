#!/usr/bin/env python
"""bench.py -- level-1 compress / decompress throughput of the B200 path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch with the input resident
in HBM; `value` is whole-job compress GB/s of UNCOMPRESSED bytes; `e2e` is the same metric through the
reference-facing C ABI call (stenos_compress_generic) on pinned HOST buffers, copies inside the timed region.

N = 1: BASELINE.json configs[1] -- 1 GiB int32 (noisy ramp + runs), level 1.  The line also carries `rows` for the
other configs of SURVEY.md section 8(d): the codec on T = 2 (int16 sine) and T = 8 (int64 ramp + runs) at 1 GiB,
the filters on the 4 GiB float64 / float32 / int16 series (config 3), the 2^20-bucket gather (config 5), the batched
bucket encode (cvector write side) and the hybrid level-3 path (device filters + host Zstd) next to the reference; and
`stream_parity`: superblocks of the very stream that was timed, compared byte for byte with the CPU oracle.

N > 1 (torchrun, one process per GPU): BASELINE.json configs[3] -- 16 GiB int16 + 16 GiB int64, each its own
frame, partitioned over the N GPUs on superblock boundaries (strong scaling).  Nothing is exchanged inside the
timed step; the segment byte lengths (8 bytes per rank and frame) are exchanged once after it.  `weak_row` keeps
the 1 GiB-per-GPU int32 line of round 1.

--impl reference times the UNTOUCHED reference (oracle/_ref/libstenos_ref.so, AVX2, all host threads) on the same
workload (N = 1: the full 1 GiB array; N > 1: a bounded 2 x 1 GiB sample of config 4), on rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "level-1 compress GB/s of uncompressed bytes, whole job (decompress GB/s, ratio and %HBM roofline in extra keys)"
WORKLOAD = "int32_ramp_runs"
T = 4
SHARD_ELEMS = 1 << 28  # 1 GiB of int32 per GPU
CFG4 = (("int16_sine", 2, 16 << 30), ("int64_ramp_runs", 8, 16 << 30))  # config 4: two frames
SB = 131072


def workload_name(world, nbytes):
    if world == 1:
        return "%d GiB int32 noisy ramp + runs (BASELINE.json configs[1]), level 1, one frame" % (nbytes >> 30) if nbytes >= (1 << 30) \
            else "%d MiB int32 noisy ramp + runs, level 1" % (nbytes >> 20)
    return "32 GiB mixed buffer (BASELINE.json configs[3]): 16 GiB int16 sine + 16 GiB int64 noisy ramp + runs, two frames, level 1, " \
           "chunk-partitioned over %d B200" % world


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


# ---------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (oracle/_ref): the baseline arm
# ---------------------------------------------------------------------------------------------------
def cpu_reference_arm(name, Tn, elems, threads, steps, warmup):
    """GB/s compress / decompress of the untouched reference on `threads` host threads."""
    from oracle import ref
    from stenos_b200 import synth

    L = ref.lib()
    a = synth.make(name, elems)
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
        r = L.stenos_compress_generic(ctx, src.ctypes.data, Tn, n, dst.ctypes.data, cap)
        dt = time.perf_counter() - t0
        if i >= warmup:
            ctimes.append(dt)
    assert not ref.has_error(r)
    # decode: the reference decoder rejects exact multiples of the superblock size: one block fewer
    n2 = n - Tn * 256
    r2 = L.stenos_compress_generic(ctx, src.ctypes.data, Tn, n2, dst.ctypes.data, cap)
    out = np.empty(n2, dtype=np.uint8)
    dtimes = []
    d = 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        d = L.stenos_decompress_generic(ctx, dst.ctypes.data, Tn, r2, out.ctypes.data, n2)
        dt = time.perf_counter() - t0
        if i >= warmup:
            dtimes.append(dt)
    assert d == n2 and out.tobytes() == src[:n2].tobytes()
    L.stenos_destroy_context(ctx)
    return {"compress_s": float(np.mean(ctimes)), "decompress_s": float(np.mean(dtimes)), "bytes": n, "csize": int(r), "ratio": n / r}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libstenos_ref.so was not built in the container"}))
        return
    world = args.gpus
    if world == 1:
        parts = [cpu_reference_arm(WORKLOAD, T, args.shard_elems, cores, args.steps, args.warmup)]
        sample = "the full %d MiB array of the workload; stenos_set_threads=%d, AVX2 build of /root/reference" % (parts[0]["bytes"] >> 20, cores)
    else:
        parts = [cpu_reference_arm(n, t, (1 << 30) // t, cores, args.steps, args.warmup) for n, t, _ in CFG4]
        sample = "1 GiB of each of the two frames of config 4 (the arrays' first GiB); stenos_set_threads=%d, AVX2 build of /root/reference" % cores
    nbytes = sum(p["bytes"] for p in parts)
    cs, ds = sum(p["compress_s"] for p in parts), sum(p["decompress_s"] for p in parts)
    line = {
        "impl": "reference", "metric": METRIC, "value": nbytes / cs / 1e9, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cs * 1e3, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(world, args.shard_elems * T), "bytesoftype": T if world == 1 else [2, 8], "superblock": SB},
        "decompress_GBps": nbytes / ds / 1e9, "ratio": nbytes / sum(p["csize"] for p in parts),
        "cpu_baseline": {"value": nbytes / cs / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": nbytes / cs / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ---------------------------------------------------------------------------------------------------
def timed_events(fn, steps, warmup, stream, sync):
    """W warm-up steps, then K steps; returns (total ms, per-step ms list) from CUDA events on `stream`."""
    import torch

    for _ in range(warmup):
        fn()
    sync()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        e0.record(stream)
        fn()
        e1.record(stream)
    sync()
    return evs[0][0].elapsed_time(evs[-1][1]), [e0.elapsed_time(e1) for e0, e1 in evs]


def stream_parity(d_src, d_dst, d_off, n_sb, Tn, nbytes, sb, k=32, seed=1, base=0):
    """k superblocks of the stream that was just timed, against the CPU oracle on the device's own input bytes.
    d_off: superblock header offsets relative to d_dst (+ base).  Returns (ok, checked)."""
    from oracle import port

    rng = np.random.RandomState(seed)
    picks = sorted(set([0, n_sb - 1] + [int(x) for x in rng.randint(0, n_sb, size=k)]))
    offs = d_off[: n_sb + 1].cpu().numpy().astype(np.int64) - base
    ok = True
    for s in picks:
        lo, hi = s * sb, min((s + 1) * sb, nbytes)
        raw = d_src[lo:hi].cpu().numpy()
        got = d_dst[int(offs[s]):int(offs[s + 1])].cpu().numpy().tobytes()
        want = port.compress_superblock(raw, Tn, room=hi - lo + 4096 + 8 * Tn)
        ok = ok and (got == want)
    return ok, len(picks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--shard-elems", type=int, default=SHARD_ELEMS)
    ap.add_argument("--cfg4-gib", type=int, default=16, help="GiB per frame of config 4 (N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rows", action="store_true")
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
    stream = torch.cuda.current_stream()
    peak, peak_src = peaks()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    class Shard:
        """One rank's segment of one frame: device buffers + the calls of a step."""

        def __init__(self, name, Tn, frame_bytes, d_src=None):
            self.name, self.T, self.frame_bytes = name, Tn, frame_bytes
            self.ctx = api.Context(level=1, stream=stream)
            self.seg = distributed.SegmentCodec(self.ctx, Tn, frame_bytes)
            self.sb = self.seg.sb
            lo, nsb, b0, cnt = distributed.plan_partition(frame_bytes, self.sb, world)[rank]
            self.first_sb, self.n_sb, self.byte0, self.nbytes = lo, nsb, b0, cnt
            self.d_src = d_src if d_src is not None else synth.make_torch(name, cnt // Tn, start=b0 // Tn, device=dev).view(torch.uint8)
            # worst case of the segment plus 1 MiB: every superblock provably has the room in which the reference's dst-room
            # checks are inert, so the whole segment is one launch of the fast encoder
            self.cap = self.seg.capacity(cnt) + (1 << 20)
            self.d_dst = torch.empty(self.cap, dtype=torch.uint8, device=dev)
            self.d_res = torch.zeros(2, dtype=torch.int64, device=dev)
            self.d_off = torch.zeros(nsb + 1, dtype=torch.int64, device=dev)
            self.csize = None

        def compress(self):
            if world == 1:
                self.ctx.compress_async(self.d_src, self.T, self.nbytes, self.d_dst, self.cap, self.d_res, self.d_off)
            else:
                self.seg.compress_async(self.d_src, self.nbytes, self.d_dst, self.cap, self.d_res, self.d_off)

        def decompress(self, d_out, with_index):
            if world == 1:
                self.ctx.decompress_async(self.d_dst, self.T, self.csize, d_out, self.nbytes, self.nbytes, self.d_res, self.d_off if with_index else None)
            else:
                self.seg.decompress_async(self.d_dst, self.csize, self.nbytes, self.d_off, d_out, self.d_res)

        def fetch_result(self):
            res = self.d_res.cpu().numpy()
            assert res[1] == 0, "device error bits %d" % res[1]
            self.csize = int(res[0])
            return self.csize

    # ---- the workload of this N
    n = args.shard_elems
    host = None
    if world == 1:
        a = synth.make(WORKLOAD, n)
        host = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
        shards = [Shard(WORKLOAD, T, n * T, d_src=host.to(dev))]
    else:
        shards = [Shard(name, t, args.cfg4_gib << 30) for name, t, _ in CFG4]
    job_bytes = sum(s.frame_bytes for s in shards)
    my_bytes = sum(s.nbytes for s in shards)
    d_out = torch.empty(max(s.nbytes for s in shards), dtype=torch.uint8, device=dev)

    def compress_step():
        for s in shards:
            s.compress()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = api.kernel_launches()
    c_ms, c_per = timed_events(compress_step, args.steps, args.warmup, stream, sync_all)
    c_launch = api.kernel_launches() - l0
    c_ms = max_over_ranks(c_ms)
    csizes = [s.fetch_result() for s in shards]

    # decompress: every frame with the superblock index the encoder produced, and (N = 1) walking the frame headers
    def decomp_step(with_index):
        for s in shards:
            s.decompress(d_out[: s.nbytes], with_index)

    l0 = api.kernel_launches()
    d_ms, d_per = timed_events(lambda: decomp_step(True), args.steps, args.warmup, stream, sync_all)
    d_launch = api.kernel_launches() - l0
    d_ms = max_over_ranks(d_ms)
    assert shards[-1].d_res.cpu().numpy()[1] == 0
    assert torch.equal(d_out[: shards[-1].nbytes], shards[-1].d_src), "round trip mismatch"
    dw_ms = None
    if world == 1:
        dw_ms, dw_per = timed_events(lambda: decomp_step(False), args.steps, args.warmup, stream, sync_all)
        assert torch.equal(d_out[: shards[0].nbytes], shards[0].d_src), "round trip mismatch (header walk)"
    clocks = sampler.stop() if rank == 0 else None
    if len(shards) > 1:
        shards[0].decompress(d_out[: shards[0].nbytes], True)
        torch.cuda.synchronize()
        assert torch.equal(d_out[: shards[0].nbytes], shards[0].d_src), "round trip mismatch"

    # ---- parity of the stream that was timed: sampled superblocks against the CPU oracle (every rank, every frame)
    par_ok, par_n = True, 0
    for s in shards:
        base = int(s.d_off[0].item())  # frames start after their header; segments at 0
        ok, k = stream_parity(s.d_src, s.d_dst, s.d_off, s.n_sb, s.T, s.nbytes, s.sb, k=32 if world == 1 else 12, seed=rank + 1, base=0)
        par_ok, par_n = par_ok and ok, par_n + k
        if world == 1:
            assert base in (8, 12)
    par_ok = all_ok(par_ok)

    # ---- N > 1: the only exchange of the job, after the step: 8 bytes per rank and frame -> frame offsets of the segments
    frame_info = None
    if world > 1:
        sizes = torch.tensor(csizes, dtype=torch.int64, device=dev)
        allsz = torch.empty(world * len(shards), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allsz, sizes)
        allsz = allsz.cpu().numpy().reshape(world, len(shards))
        frame_info = []
        for j, s in enumerate(shards):
            hdr = distributed.frame_header(s.frame_bytes)
            starts = len(hdr) + np.concatenate([[0], np.cumsum(allsz[:, j])[:-1]])
            frame_info.append({"frame": s.name, "bytes": s.frame_bytes, "compressed_bytes": int(len(hdr) + allsz[:, j].sum()),
                               "segment_offsets": [int(x) for x in starts], "ratio": s.frame_bytes / float(len(hdr) + allsz[:, j].sum())})
        # the assembled frame [header] + segments decodes like a single-GPU stream: every rank's segment starts where the
        # exclusive prefix says, and superblock s of the frame is superblock s - first_sb of its rank's segment (checked
        # above against the oracle byte for byte)

    # ---- weak row at N > 1: 1 GiB int32 per GPU (round 1's line), nothing exchanged in the step either
    weak = None
    if world > 1:
        w = Shard(WORKLOAD, T, n * T * world)
        w_ms, _ = timed_events(w.compress, args.steps, args.warmup, stream, sync_all)
        w_ms = max_over_ranks(w_ms)
        w.fetch_result()
        okw, kw = stream_parity(w.d_src, w.d_dst, w.d_off, w.n_sb, T, w.nbytes, w.sb, k=8, seed=rank + 7)
        weak = {"workload": "1 GiB int32 noisy ramp + runs per GPU, one frame of %d GiB" % world, "scaling": "weak", "compress_GBps": n * T * world / (w_ms / args.steps * 1e-3) / 1e9,
                "ms_per_step": w_ms / args.steps, "stream_parity": all_ok(okw)}
        del w

    # ---- e2e: the reference-facing calls on pinned host buffers (H2D + kernels + D2H inside the timed region), next to
    # the plain pinned copies of the same bytes on the same ranks at the same time (the PCIe ceiling of this box)
    e2e = None
    if not args.no_e2e:
        if host is None:
            host = torch.from_numpy(np.ascontiguousarray(synth.make(WORKLOAD, n, start=rank * n)).view(np.uint8).reshape(-1))
        nb = host.numel()
        h_src = host.pin_memory()
        h_dst = torch.empty(api.bound(nb), dtype=torch.uint8).pin_memory()
        h_back = torch.empty(nb, dtype=torch.uint8).pin_memory()
        ectx = api.Context(level=1, stream=stream)
        esteps = max(3, min(args.steps, 5))

        def wall(fn):
            for _ in range(2):
                fn()
            sync_all()
            t0 = time.perf_counter()
            for _ in range(esteps):
                fn()
            sync_all()
            return max_over_ranks((time.perf_counter() - t0) / esteps)

        r = [0]

        def e_comp():
            r[0] = api.check(ectx.compress_raw(h_src, T, nb, h_dst, h_dst.numel()), "stenos_compress_generic")

        def e_decomp():
            api.check(ectx.decompress_raw(h_dst, T, r[0], h_back, nb), "stenos_decompress_generic")

        d_tmp = torch.empty(nb, dtype=torch.uint8, device=dev)
        t_c = wall(e_comp)
        t_d = wall(e_decomp)
        assert torch.equal(h_back, h_src), "e2e round trip mismatch"
        t_h2d = wall(lambda: (d_tmp.copy_(h_src, non_blocking=True), torch.cuda.synchronize()))
        t_d2h = wall(lambda: (h_back.copy_(d_tmp, non_blocking=True), torch.cuda.synchronize()))
        e2e = {"value": nb * world / t_c / 1e9, "unit": "GB/s", "h2d_bytes_per_step": nb, "d2h_bytes_per_step": int(r[0]) + 16,
               "call": "stenos_compress_generic(ctx, pinned host src, 4, bytes, pinned host dst, stenos_bound(bytes)); %d MiB int32 per rank" % (nb >> 20),
               "decompress_GBps": nb * world / t_d / 1e9, "decompress_call": "stenos_decompress_generic(ctx, pinned host frame, 4, csize, pinned host dst, bytes)",
               "decompress_h2d_bytes_per_step": int(r[0]), "decompress_d2h_bytes_per_step": nb,
               "pinned_h2d_ceiling_GBps": nb * world / t_h2d / 1e9, "pinned_d2h_ceiling_GBps": nb * world / t_d2h / 1e9,
               "frac_of_h2d_ceiling": t_h2d / t_c, "decompress_frac_of_d2h_ceiling": t_d2h / t_d}
        ectx.close()
        del d_tmp

    # ---- rows of the other configs (N = 1 only)
    rows = []
    if world == 1 and not args.no_rows:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_rows as rowlib

        rows = rowlib.all_rows(dev, stream, peak, steps=max(3, min(args.steps, 5)), warmup=3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    c_step_ms = c_ms / args.steps
    d_step_ms = d_ms / args.steps
    k_ms = float(np.mean(c_per))  # device time of one compress step on the launching stream (the memsets of the control block + the encode kernel)
    alg_bytes = my_bytes + sum(csizes)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("encode_flow_kernel_dram_bytes_per_launch") if world == 1 else None
            dtraffic = tj.get("decode_pairs_kernel_dram_bytes_per_launch") if world == 1 else None
        except Exception:
            traffic = dtraffic = None
    else:
        dtraffic = None
    line = {
        "metric": METRIC, "value": job_bytes / (c_step_ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": c_step_ms, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(world, n * T), "bytesoftype": T if world == 1 else [2, 8], "superblock": SB,
                   "l2": "inputs larger than L2 (no flush needed)", "device_dst_capacity": "worst case of the segment + 1 MiB",
                   "parallelism": "superblock ranges per GPU; segment sizes exchanged once after the step (8 bytes per rank and frame)" if world > 1 else "1 GPU"},
        "decompress_GBps": job_bytes / (d_step_ms * 1e-3) / 1e9, "decompress_ms_per_step": d_step_ms,
        "decompress_with_header_walk_GBps": (job_bytes / (dw_ms / args.steps * 1e-3) / 1e9) if dw_ms else None,
        "ratio": my_bytes / float(sum(csizes)), "compressed_bytes_per_gpu": int(sum(csizes)),
        "stream_parity": par_ok, "stream_parity_superblocks": par_n,
        "roofline": {"bound": "hbm", "kernel": "encode_flow_kernel<4,576,4>" if world == 1 else "encode_flow_kernel<2,...> + <8,...> (one launch per frame)",
                     "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms},
        "roofline_decompress": {"bound": "hbm", "kernel": "decode_pairs_kernel (with the encoder's superblock index)", "achieved": alg_bytes / (float(np.mean(d_per)) * 1e-3) / 1e9,
                                "peak": peak, "unit": "GB/s", "frac": alg_bytes / (float(np.mean(d_per)) * 1e-3) / 1e9 / peak, "traffic": dtraffic},
        "gpu_launches": int(c_launch + d_launch), "clocks": clocks,
    }
    if frame_info:
        line["frames"] = frame_info
    if weak:
        line["weak_row"] = weak
    if e2e:
        line["e2e"] = e2e
    if rows:
        line["rows"] = rows
    if not args.no_cpu_baseline and world == 1:
        from oracle import ref
        if ref.available():
            cores = os.cpu_count() or 1
            m = cpu_reference_arm(WORKLOAD, T, n, cores, 3, 1)
            m1 = cpu_reference_arm(WORKLOAD, T, 1 << 24, 1, 3, 1)
            line["cpu_baseline"] = {"value": m["bytes"] / m["compress_s"] / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference",
                                    "sample": "the full %d MiB array; stenos_set_threads=%d" % (m["bytes"] >> 20, cores),
                                    "decompress_GBps": m["bytes"] / m["decompress_s"] / 1e9, "single_thread_compress_GBps": m1["bytes"] / m1["compress_s"] / 1e9,
                                    "single_thread_decompress_GBps": m1["bytes"] / m1["decompress_s"] / 1e9, "ratio": m["ratio"],
                                    "ratio_identical_to_gpu": abs(m["ratio"] - line["ratio"]) < 1e-12 if m["bytes"] == my_bytes else None}
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
