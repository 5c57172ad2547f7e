#!/usr/bin/env python
"""Attributes the per-instruction counters of an `ncu --set full --import-source on` report to CUDA
source lines.  ncu's CSV source page is SASS only; line numbers come from `nvdisasm -g` on the cubin
of the same build (instructions are matched by position inside the kernel).

    python tools/ncu_lines.py gpurun_out/enc_r1.ncu-rep 'encode_frame_kernelILi4' [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel_substr):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "stenos_b200", "libstenos_b200.so")], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    out, cur, inside = [], None, False
    for line in txt.splitlines():
        if line.startswith("//--------------------- .text."):
            inside = kernel_substr in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            out.append(cur)
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hi]
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    inst = [(int(r[ii] or 0), int(r[si] or 0), r[1]) for r in rows[hi + 1:] if len(r) > ii]
    lines = sass_lines(kern)
    if len(lines) != len(inst):
        print("warning: %d SASS instructions in the cubin vs %d in the report" % (len(lines), len(inst)))
    agg = {}
    for (n, s, _), ln in zip(inst, lines):
        a = agg.setdefault(ln, [0, 0, 0])
        a[0] += n
        a[1] += s
        a[2] += 1
    tot, tots = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print("total warp instructions %d, stall samples %d" % (tot, tots))
    src_cache = {}
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if ln:
            p = os.path.join(ROOT, "stenos_b200", "csrc", ln[0])
            if os.path.exists(p):
                src_cache.setdefault(p, open(p).read().splitlines())
                if ln[1] - 1 < len(src_cache[p]):
                    text = src_cache[p][ln[1] - 1].strip()[:90]
        print("%10d %5.1f%%  samples %5.1f%%  sass %4d  %s:%s  %s" % (a[0], 100.0 * a[0] / tot, 100.0 * a[1] / max(tots, 1), a[2], ln[0] if ln else "?", ln[1] if ln else "?", text))


if __name__ == "__main__":
    main()
