#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_f16x3.py tests/test_gpu_bf16x3.py tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -4
for m in f16x3 bf16; do
DPN_PHASE_DEBUG=1 timeout 120 python bench.py --mode $m --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes 2>&1 >/dev/null | grep "dpn phase" | tail -2
timeout 120 python tools/step_jitter.py $m 20 2>&1 | grep -E "per-step"
done
