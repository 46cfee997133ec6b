#!/bin/bash
# A/B of a kernel-variant build against the in-tree library: DPN_LIB_OVERRIDE=<variant .so>.  usage: gpu_ab.sh <variant.so> [modes]
cd "$(dirname "$0")/.."
VAR=$1; MODES=${2:-"f16x3 bf16"}
for lib in "" "$VAR"; do
  echo "== library: ${lib:-in-tree}"
  for m in $MODES; do
    DPN_LIB_OVERRIDE=$lib timeout 90 python tools/step_jitter.py $m 16 2>&1 | grep -E "per-step" | cut -c1-170
  done
done
DPN_LIB_OVERRIDE=$VAR timeout 150 python -m pytest tests/test_gpu_f16x3.py -m gpu -x -q 2>&1 | tail -2
