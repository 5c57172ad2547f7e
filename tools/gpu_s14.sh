#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/bench_rows.py --rows codec > gpurun_out/r2s14_rows_flow.jsonl 2> gpurun_out/r2s14_rows_flow.err
for v in vR15 vR11 vD1 vD2 vD3; do
STENOS_B200_LIB=build/variants/$v.so timeout 150 python tools/bench_rows.py --rows codec > gpurun_out/r2s14_rows_$v.jsonl 2> gpurun_out/r2s14_rows_$v.err
done
for f in gpurun_out/r2s14_rows_*.jsonl; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("  T=%d compress %.3f ms frac %.3f | decompress %.3f ms frac %.3f (walk %.3f) parity %s"%(d["T"],d["compress_ms"],d["compress_frac"],d["decompress_ms"],d["decompress_frac"],d["decompress_with_header_walk_ms"],d.get("parity")))
PY
done
