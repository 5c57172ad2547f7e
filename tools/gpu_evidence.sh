#!/bin/bash
# ncu evidence of the round's kernels, summarised ON the GPU box (gpurun brings back at most 64 MiB: the reports stay in /tmp there,
# except the T=4 encoder's): full captures of one launch of encode_flow_kernel (T = 4, 2, 8) and decode_pairs_kernel (T = 4), and the
# launch list of the bench command.   usage: gpu_evidence.sh <tag>
R=${1:-r02}
mkdir -p gpurun_out /tmp/ncu
for T in 4 2 8; do
TP_T=$T timeout 300 ncu --set full --import-source on --clock-control none -k regex:encode_flow_kernel -c 1 -s 3 -o /tmp/ncu/${R}_enc$T -f python tools/time_parts.py > gpurun_out/${R}_ncu_enc$T.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${R}_enc$T.ncu-rep encode_flow_kernelILi$T > gpurun_out/${R}_ncu_encode_flow_T$T.txt 2>&1
done
TP_T=4 timeout 300 ncu --set full --import-source on --clock-control none -k regex:decode_pairs_kernel -c 1 -s 3 -o /tmp/ncu/${R}_dec4 -f python tools/time_parts.py > gpurun_out/${R}_ncu_dec4.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${R}_dec4.ncu-rep decode_pairs_kernelILi4 > gpurun_out/${R}_ncu_decode_pairs_T4.txt 2>&1
cp /tmp/ncu/${R}_enc4.ncu-rep gpurun_out/
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/${R}_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12
