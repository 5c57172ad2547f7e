#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "hybrid or bucket" > gpurun_out/b5_pytest.log 2>&1; tail -3 gpurun_out/b5_pytest.log
timeout 300 python tools/bench_rows.py --rows buckets,gather > gpurun_out/b5_rows.jsonl 2> gpurun_out/b5_rows.err; cut -c1-500 gpurun_out/b5_rows.jsonl; tail -3 gpurun_out/b5_rows.err
