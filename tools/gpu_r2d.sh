#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "filters or index" > gpurun_out/b3_pytest.log 2>&1; tail -2 gpurun_out/b3_pytest.log
timeout 400 python tools/bench_rows.py --rows filters > gpurun_out/b3_rows_filters.jsonl 2> gpurun_out/b3_rows_filters.err; cut -c1-420 gpurun_out/b3_rows_filters.jsonl
tools/gpu_r2c.sh b3 "4 2" main span32
