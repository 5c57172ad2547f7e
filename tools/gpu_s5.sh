#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_rows.py --rows codec > gpurun_out/r2s6_rows_flow.jsonl 2> gpurun_out/r2s6_rows_flow.err
for v in ; do
STENOS_B200_LIB=build/variants/$v.so timeout 600 python tools/bench_rows.py --rows codec > gpurun_out/r2s6_rows_$v.jsonl 2> gpurun_out/r2s6_rows_$v.err
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o gpurun_out/r2s6_enc4 -f python tools/time_parts.py > gpurun_out/r2s6_ncu_enc4.log 2>&1
STENOS_B200_LIB=build/variants/vS0.so timeout 900 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o gpurun_out/r2s6_enc4_vS0 -f python tools/time_parts.py > gpurun_out/r2s6_ncu_enc4_vS0.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2s6_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2s6_pytest.log
tail -3 gpurun_out/r2s6_pytest.log
for f in gpurun_out/r2s6_rows_*.jsonl; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("  T=%d compress %.3f ms frac %.3f | decompress %.3f ms frac %.3f"%(d["T"],d["compress_ms"],d["compress_frac"],d["decompress_ms"],d["decompress_frac"]))
PY
done
