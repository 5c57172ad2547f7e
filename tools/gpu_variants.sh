#!/bin/bash
# variants of the library timed with tools/time_parts.py: gpu_variants.sh <tag> <T list> <variant names ...> ("main" = the in-tree library)
TAG=$1; TS=$2; shift 2; mkdir -p gpurun_out
for V in "$@"; do
  for T in $TS; do
    if [ "$V" = main ]; then L=""; else L="build/variants/$V.so"; fi
    echo -n "$V T=$T: " | tee -a gpurun_out/${TAG}_parts.log
    STENOS_B200_LIB=$L TP_T=$T timeout 200 python tools/time_parts.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_parts.log
  done
done
