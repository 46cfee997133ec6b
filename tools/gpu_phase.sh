#!/bin/bash
# quick kernel iteration loop: split-mode parity tests + per-step time + DPN_PHASE_DEBUG cycle counters
cd "$(dirname "$0")/.."
timeout 200 python -m pytest tests/test_gpu_f16x3.py tests/test_gpu_bf16x3.py tests/test_gpu_bf16.py tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -3
for m in ${MODES:-f16x3 bf16}; do
timeout 90 python tools/step_jitter.py $m 16 2>&1 | grep -E "per-step" | cut -c1-170
DPN_PHASE_DEBUG=1 timeout 90 python tools/step_jitter.py $m 4 > /tmp/sj.txt 2>&1
grep -E "phase. pass1" /tmp/sj.txt | tail -1 | cut -c1-330
grep -E "phase. pass2" /tmp/sj.txt | tail -1 | cut -c1-330
done
