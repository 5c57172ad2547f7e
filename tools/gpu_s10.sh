#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r2s10_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2s10_pytest.log
tail -5 gpurun_out/r2s10_pytest.log
timeout 150 python tools/bench_rows.py --rows codec > gpurun_out/r2s10_rows_flow.jsonl 2> gpurun_out/r2s10_rows_flow.err
for v in vW18 vW19; do
STENOS_B200_LIB=build/variants/$v.so timeout 150 python tools/bench_rows.py --rows codec > gpurun_out/r2s10_rows_$v.jsonl 2> gpurun_out/r2s10_rows_$v.err
done
for f in gpurun_out/r2s10_rows_*.jsonl; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("  T=%d compress %.3f ms frac %.3f | decompress %.3f ms frac %.3f parity %s"%(d["T"],d["compress_ms"],d["compress_frac"],d["decompress_ms"],d["decompress_frac"],d.get("parity")))
PY
done
timeout 420 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s10_bench.json 2> gpurun_out/r2s10_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r2s10_bench.err; cut -c1-1500 gpurun_out/r2s10_bench.json
