#!/bin/bash
tools/gpu_final_n1.sh r02g
NOCHECK=1 tools/gpu_r2c.sh b10 "4 2" main da1 da3 da4
