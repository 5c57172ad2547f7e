#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "forced or hybrid" > gpurun_out/b6_pytest.log 2>&1; tail -3 gpurun_out/b6_pytest.log
timeout 400 python tools/time_hybrid.py --mib 256 --level 3 > gpurun_out/b6_hybrid.jsonl 2> gpurun_out/b6_hybrid.err; cat gpurun_out/b6_hybrid.jsonl; tail -3 gpurun_out/b6_hybrid.err
