#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_f16x3.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
timeout 200 python -m pytest tests/test_gpu_bf16x3.py tests/test_gpu_bf16.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --mode f16x3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-modes 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('f16x3 value %.2f Mpts/s %.2f ms' % (d['value']/1e6, d['ms_per_step']))"
