#!/bin/bash
cd "$(dirname "$0")/.."
for f in 8 0; do
echo "== DPN_DEBUG_FLAGS=$f (8 = no L2 prefetch)"
DPN_DEBUG_FLAGS=$f DPN_PHASE_DEBUG=1 timeout 120 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes 2>&1 >/dev/null | grep "dpn phase" | grep pass2 | tail -1
DPN_DEBUG_FLAGS=$f timeout 120 python tools/step_jitter.py bf16 30 2>&1 | grep -E "per-step"
done
timeout 200 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q 2>&1 | tail -2
