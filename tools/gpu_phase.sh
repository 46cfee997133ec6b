#!/bin/bash
cd "$(dirname "$0")/.."
timeout 150 python -m pytest tests/test_gpu_f16x3.py tests/test_gpu_bf16.py -m gpu -x -q 2>&1 | tail -3
for m in f16x3 bf16; do
DPN_PHASE_DEBUG=1 timeout 90 python tools/step_jitter.py $m 12 > /tmp/sj.txt 2>&1
grep -E "per-step" /tmp/sj.txt | cut -c1-200
grep -E "phase. pass1" /tmp/sj.txt | tail -1 | cut -c1-330
grep -E "phase. pass2" /tmp/sj.txt | tail -1 | cut -c1-330
done
