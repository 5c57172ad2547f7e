#!/bin/bash
# ncu evidence of the round's kernels: full captures (one launch each) + the launch list of the bench command
# usage: gpu_evidence.sh <tag> <part: a|b>   (two calls: gpurun brings back at most 64 MiB)
R=${1:-r02}; PART=${2:-a}
mkdir -p gpurun_out
if [ "$PART" = a ]; then
TP_T=4 timeout 300 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o gpurun_out/${R}_enc4 -f python tools/time_parts.py > gpurun_out/${R}_ncu_enc4.log 2>&1
TP_T=4 timeout 300 ncu --set full --import-source on --clock-control none -k regex:decode_pairs_kernel -c 1 -s 3 -o gpurun_out/${R}_dec4 -f python tools/time_parts.py > gpurun_out/${R}_ncu_dec4.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/${R}_bench_under_ncu.log 2>&1
else
for T in 2 8; do
TP_T=$T timeout 300 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o gpurun_out/${R}_enc$T -f python tools/time_parts.py > gpurun_out/${R}_ncu_enc$T.log 2>&1
done
fi
ls -la gpurun_out | tail -8
