#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2s13_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2s13_bench_n2.json 2> gpurun_out/r2s13_bench_n2.err; echo "bench n2 rc $?"
tail -5 gpurun_out/r2s13_bench_n2.err; cut -c1-3000 gpurun_out/r2s13_bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2s13_ref_n2.json 2> gpurun_out/r2s13_ref_n2.err; echo "ref n2 rc $?"
tail -3 gpurun_out/r2s13_ref_n2.err; cut -c1-1200 gpurun_out/r2s13_ref_n2.json
