#!/bin/bash
# second batch of ncu summaries (on the box): decode_pairs_kernel T = 2 / 8, unshuffle_delta_kernel<8>, gather_pairs_kernel<4>, encode_bucket_pairs_kernel<4>
R=${1:-r02}; mkdir -p gpurun_out /tmp/ncu
for T in 2 8; do
TP_T=$T timeout 300 ncu --set full --import-source on --clock-control none -k regex:decode_pairs_kernel -c 1 -s 3 -o /tmp/ncu/${R}_dec$T -f python tools/time_parts.py > gpurun_out/${R}_ncu_dec$T.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${R}_dec$T.ncu-rep decode_pairs_kernelILi$T > gpurun_out/${R}_ncu_decode_pairs_T$T.txt 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:unshuffle_delta_kernel -c 1 -s 2 -o /tmp/ncu/${R}_ud8 -f python tools/bench_rows.py --rows filters8 --filter-mib 1024 > gpurun_out/${R}_ncu_ud8.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${R}_ud8.ncu-rep unshuffle_delta_kernelILi8 > gpurun_out/${R}_ncu_unshuffle_delta_T8.txt 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gather_pairs_kernel -c 1 -s 2 -o /tmp/ncu/${R}_gather -f python tools/bench_rows.py --rows gather > gpurun_out/${R}_ncu_gather.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${R}_gather.ncu-rep gather_pairs_kernelILi4 > gpurun_out/${R}_ncu_gather_pairs_T4.txt 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:encode_bucket_pairs_kernel -c 1 -s 2 -o /tmp/ncu/${R}_bk -f python tools/bench_rows.py --rows buckets > gpurun_out/${R}_ncu_bk.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${R}_bk.ncu-rep encode_bucket_pairs_kernelILi4 > gpurun_out/${R}_ncu_encode_bucket_pairs_T4.txt 2>&1
for f in gpurun_out/${R}_ncu_decode_pairs_T2.txt gpurun_out/${R}_ncu_decode_pairs_T8.txt gpurun_out/${R}_ncu_unshuffle_delta_T8.txt gpurun_out/${R}_ncu_gather_pairs_T4.txt gpurun_out/${R}_ncu_encode_bucket_pairs_T4.txt; do echo "== $f"; sed -n 1,30p $f | grep -E "duration|registers|warps_active|inst_executed.sum|issue_active|pipe_alu|dram__bytes|long_scoreboard|wait  "; done
