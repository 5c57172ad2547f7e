#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_rows.py --rows codec > gpurun_out/r2s3_rows_flow.jsonl 2> gpurun_out/r2s3_rows_flow.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o gpurun_out/r2s3_enc4 -f python tools/time_parts.py > gpurun_out/r2s3_ncu_enc4.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2s3_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2s3_pytest.log
tail -3 gpurun_out/r2s3_pytest.log; cat gpurun_out/r2s3_rows_flow.jsonl | cut -c1-330
