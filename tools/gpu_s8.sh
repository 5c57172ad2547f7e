#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_rows.py --rows codec > gpurun_out/r2s9_rows_flow.jsonl 2> gpurun_out/r2s9_rows_flow.err
for v in vW18 vW19; do
STENOS_B200_LIB=build/variants/$v.so timeout 600 python tools/bench_rows.py --rows codec > gpurun_out/r2s9_rows_$v.jsonl 2> gpurun_out/r2s9_rows_$v.err
done
STENOS_B200_LIB=build/variants/vW19.so timeout 900 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o gpurun_out/r2s9_enc4_w19 -f python tools/time_parts.py > gpurun_out/r2s9_ncu_enc4.log 2>&1
STENOS_B200_LIB=build/variants/vW19.so TP_T=2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o gpurun_out/r2s9_enc2_w19 -f python tools/time_parts.py > gpurun_out/r2s9_ncu_enc2.log 2>&1
for f in gpurun_out/r2s9_rows_*.jsonl; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("  T=%d compress %.3f ms frac %.3f | decompress %.3f ms frac %.3f"%(d["T"],d["compress_ms"],d["compress_frac"],d["decompress_ms"],d["decompress_frac"]))
PY
done
