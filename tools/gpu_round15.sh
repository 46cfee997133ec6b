#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_bf16x3.py -m gpu -x -q 2>&1 | tail -15
timeout 200 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --mode bf16x3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-modes 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bf16x3 value %.2f Mpts/s %.2f ms' % (d['value']/1e6, d['ms_per_step']))"
