#!/bin/bash
# session-2 iteration script: GPU suite (optional) + per-T timings of the encoder / decoder
# usage: gpu_r2b.sh <tag> [tests]
TAG=${1:-b1}; mkdir -p gpurun_out
if [ "$2" = tests ]; then
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
fi
for T in 4 2 8; do TP_T=$T timeout 200 python tools/time_parts.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_parts.log; done
