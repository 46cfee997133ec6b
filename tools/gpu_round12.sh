#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,power.max_limit,clocks.max.sm,temperature.gpu --format=csv
timeout 300 python tools/step_jitter.py bf16 80 2>&1 | tail -25
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-modes 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.2f Mpts/s %.2f ms  e2e %.2f ms' % (d['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step']), d['clocks'])"
