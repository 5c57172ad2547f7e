#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b11_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b11_pytest.log; tail -3 gpurun_out/b11_pytest.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:unshuffle_delta_kernel -c 1 -s 2 -o /tmp/ncu/ud8 -f python tools/bench_rows.py --rows filters8 --filter-mib 1024 > gpurun_out/b11_ncu_ud8.log 2>&1
python tools/ncu_summary.py /tmp/ncu/ud8.ncu-rep unshuffle_delta_kernelILi8 > gpurun_out/b11_ncu_unshuffle_delta_T8.txt 2>&1
head -40 gpurun_out/b11_ncu_unshuffle_delta_T8.txt
