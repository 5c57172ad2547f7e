#!/bin/bash
# always-spill experiment: rings at their legal minimum (STENOS_B200_FLOW_RING=1): every piece goes through its HBM spill slot
mkdir -p gpurun_out
NOCHECK=1 tools/gpu_r2c.sh b9 "4" main
echo "--- rings at the minimum (always spill)" | tee -a gpurun_out/b9_parts.log
STENOS_B200_FLOW_RING=1 NOCHECK=1 tools/gpu_r2c.sh b9 "4" main k8 k16
