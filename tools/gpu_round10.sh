#!/bin/bash
cd "$(dirname "$0")/.."
for f in 0 1 2 4 7; do
echo "== DPN_DEBUG_FLAGS=$f"
DPN_DEBUG_FLAGS=$f DPN_PHASE_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes 2>&1 >/dev/null | grep "dpn phase" | grep pass1 | tail -1
done
