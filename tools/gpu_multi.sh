#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh N [tag]'
#   -> R-rank vs 1-rank gradient parity (both sharding levels), our arm at N GPUs (weak + strong scaling keys, e2e), e2e phase breakdown
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
for m in fp32 f16x3; do
  tol=1e-6; [ $m = f16x3 ] && tol=2e-5
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      tests/multirank_child.py $m $tol 2>&1 | grep -E "MULTIRANK|Error|error" | head -5 | tee -a gpurun_out/${TAG}_multirank_n${N}.txt
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --e2e-breakdown 2> gpurun_out/${TAG}_bench_n${N}_err.txt | tee gpurun_out/${TAG}_bench_n${N}.json | cut -c1-2500
grep -E "e2e phases" gpurun_out/${TAG}_bench_n${N}_err.txt | cut -c1-400 | tee gpurun_out/${TAG}_e2e_breakdown_n${N}.txt
tail -3 gpurun_out/${TAG}_bench_n${N}_err.txt | cut -c1-300
