#!/bin/bash
mkdir -p gpurun_out
NOCHECK=1 tools/gpu_r2c.sh b8 "4 8" main k5 k4 t8w15
