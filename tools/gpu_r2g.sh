#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fuzz or golden or mixed or spill or holds or full_size or edge" > gpurun_out/b7_pytest.log 2>&1; tail -3 gpurun_out/b7_pytest.log
NOCHECK=1 tools/gpu_r2c.sh b7 "4 2 8" main k6 k16 nt640
