#!/bin/bash
cd "$(dirname "$0")/.."
timeout 180 python -m pytest tests/test_gpu_f16x3.py tests/test_gpu_bf16x3.py -m gpu -x -q 2>&1 | tail -4
timeout 120 python tools/step_jitter.py f16x3 16 2>&1 | grep -E "per-step"
DPN_WGRAD_SINGLE_STAGE=1 timeout 120 python tools/step_jitter.py f16x3 16 2>&1 | grep -E "per-step"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"wgrad" -c 2 --csv --log-file gpurun_out/wg2.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > /dev/null 2>&1
grep -E "wgrad" gpurun_out/wg2.csv | cut -d, -f5,13- | head -4
