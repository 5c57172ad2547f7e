#!/usr/bin/env python
"""Compact text summary of one kernel of an `ncu --set full --import-source on` report: the headline counters, the stall
reasons per issued instruction, the executed SASS opcode histogram and the hottest source lines (tools/ncu_lines.py).

    python tools/ncu_summary.py gpurun_out/enc_s6.ncu-rep encode_stream_kernelILi4 > profiles/r01_ncu_encode.txt
"""
import collections, csv, io, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]

def main():
    rep, kern = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    d = dict(zip(h, zip(u, v)))
    print("report %s, kernel %s" % (os.path.basename(rep), d.get("Kernel Name", ("", "?"))[1]))
    for k in KEYS:
        if k in d:
            print("  %-66s %s %s" % (k, d[k][1], d[k][0]))
    print("stall reasons (warp cycles per issued instruction):")
    st = []
    for k in h:
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
        if m and not m.group(1).endswith("not_issued"):
            try:
                st.append((float(d[k][1].replace(",", "")), m.group(1)))
            except ValueError:
                pass
    for a, k in sorted(st, reverse=True)[:10]:
        print("  %-24s %.2f" % (k, a))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(srows) if "Address" in r and "Source" in r)
    hdr = srows[hi]
    ii, si = hdr.index("Instructions Executed"), hdr.index("Source")
    ops = collections.Counter()
    tot = 0
    for r in srows[hi + 1:]:
        if len(r) <= ii:
            continue
        n = int(r[ii] or 0)
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
        op = m.group(2) if m else "?"
        op = ".".join(op.split(".")[:3]) if op.startswith(("LDG", "STG", "LDS", "STS", "LDGSTS", "ATOM", "RED")) else op.split(".")[0]
        ops[op] += n
        tot += n
    print("executed warp instructions by opcode (%d total):" % tot)
    for op, n in ops.most_common(28):
        print("  %-16s %12d %5.1f%%" % (op, n, 100.0 * n / tot))
    print("hottest source lines:")
    sys.stdout.flush()
    sys.argv = ["ncu_lines.py", rep, kern, "25"]
    ncu_lines.main()

main()
