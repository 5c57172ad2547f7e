#!/bin/bash
# round-2 session 1: parity of the flow encoder on the GPU, codec rows for both encoders, ncu of the new kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2s1_pytest.log
timeout 600 python tools/bench_rows.py --rows codec > gpurun_out/r2s1_rows_flow.jsonl 2> gpurun_out/r2s1_rows_flow.err
STENOS_B200_ENCODER=1 timeout 600 python tools/bench_rows.py --rows codec > gpurun_out/r2s1_rows_v1.jsonl 2> gpurun_out/r2s1_rows_v1.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o gpurun_out/r2s1_enc4 -f python tools/time_parts.py > gpurun_out/r2s1_ncu_enc4.log 2>&1
TP_T=2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o gpurun_out/r2s1_enc2 -f python tools/time_parts.py > gpurun_out/r2s1_ncu_enc2.log 2>&1
tail -3 gpurun_out/r2s1_pytest.log; cat gpurun_out/r2s1_rows_flow.jsonl gpurun_out/r2s1_rows_v1.jsonl | cut -c1-400
