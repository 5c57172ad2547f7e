#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "filters" > gpurun_out/b12_pytest.log 2>&1; tail -2 gpurun_out/b12_pytest.log
timeout 400 python tools/bench_rows.py --rows filters --filter-mib 1024 > gpurun_out/b12_rows_filters.jsonl 2> gpurun_out/b12_rows_filters.err
python - <<'P'
import json
for l in open('gpurun_out/b12_rows_filters.jsonl'):
    r=json.loads(l); print(r['T'], r['chunk'], {k:round(v,3) for k,v in r.items() if k.endswith('_frac') or k.endswith('_ms')}, r['parity'])
P
tail -2 gpurun_out/b12_rows_filters.err
