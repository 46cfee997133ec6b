#!/bin/bash
# usage: gpu_multi.sh N  -> our arm and the reference arm at N GPUs
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_n${N}_err.txt | tee gpurun_out/bench_n${N}.json | cut -c1-1500
tail -3 gpurun_out/bench_n${N}_err.txt | cut -c1-300
