#!/bin/bash
# bench.py on N GPUs of one box, both arms: gpu_bench_n.sh <tag> <N>
TAG=$1; N=$2; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo_n$N.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench n$N rc $?"
tail -4 gpurun_out/${TAG}_bench_n$N.err; cut -c1-2500 gpurun_out/${TAG}_bench_n$N.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref_n$N.json 2> gpurun_out/${TAG}_bench_ref_n$N.err; echo "ref n$N rc $?"
cut -c1-700 gpurun_out/${TAG}_bench_ref_n$N.json
