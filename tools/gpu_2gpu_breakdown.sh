#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 10 --warmup 3 --e2e-breakdown 2> gpurun_out/b2_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'])"
grep "e2e phases" gpurun_out/b2_err.txt
