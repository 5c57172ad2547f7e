"""World-size-2 `gloo` test of the multi-GPU partition logic on CPU.

Each rank encodes its superblock range into a segment with the (emulated) kernels, the segment byte
lengths are all-gathered, rank 0 assembles [header] + segments and checks the result is the
reference's single-device stream; then each rank decodes its own range back."""
import os
import socket
import subprocess
import sys

import numpy as np

from stenos_b200 import distributed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import emu_lib; emu_lib.activate()
from stenos_b200 import api, distributed, synth
from oracle import port
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
T = 4
total_elems = (5 * 131072 + 3000) // T          # 5 full superblocks + a partial one
a = synth.make("int32_ramp_runs", total_elems)
total = a.nbytes
ctx = api.Context()
seg = distributed.SegmentCodec(ctx, T, total)
plan = distributed.plan_partition(total, seg.sb, world)
first_sb, n_sb, b0, nb = plan[rank]
src = np.ascontiguousarray(a).view(np.uint8)[b0:b0 + nb].copy()
dst = np.zeros(seg.capacity(nb), dtype=np.uint8)
res = np.zeros(2, dtype=np.uint64)
offs = np.zeros(n_sb + 1, dtype=np.uint64)
seg.compress_async(src, nb, dst, dst.size, res, offs)
ctx.synchronize()
assert res[1] == 0
sizes, base = distributed.exchange_segment_sizes(int(res[0]), dist, "cpu")
# only sizes were exchanged so far; gather the segments to rank 0 just to check the frame
parts = [None] * world
dist.all_gather_object(parts, dst[: int(res[0])].tobytes())
if rank == 0:
    frame = distributed.frame_header(total) + b"".join(parts)
    want = port.compress(a, T)
    assert frame == want, (len(frame), len(want))
    assert 8 + int(sizes.sum()) == len(want)
# decode my own range from my own segment
out = np.zeros(nb, dtype=np.uint8)
seg.decompress_async(dst, int(res[0]), nb, offs, out, res)
ctx.synchronize()
assert res[1] == 0 and out.tobytes() == src.tobytes()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", int(sizes[rank]), base)
"""


def test_plan_partition():
    plan = distributed.plan_partition(5 * 131072 + 3000, 131072, 2)
    assert plan == [(0, 3, 0, 3 * 131072), (3, 3, 3 * 131072, 2 * 131072 + 3000)]
    plan = distributed.plan_partition(1 << 35, 131072, 8)
    assert sum(p[3] for p in plan) == 1 << 35 and all(p[3] % 131072 == 0 for p in plan)
    assert distributed.plan_partition(1000, 131072, 4)[-1] == (0, 1, 0, 1000)  # fewer superblocks than ranks
    assert distributed.frame_header(1000000) == bytes([0, 0x40, 0x42, 0x0F, 0, 0, 0, 0])


def test_two_rank_gloo_partition():
    # build the emulator library once, here: two ranks running `make` on the same target at the same time race
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % {"root": ROOT}], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
