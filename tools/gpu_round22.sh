#!/bin/bash
cd "$(dirname "$0")/.."
for lib in "" tools/bin/libdpn_stage3.so; do
echo "== lib override: '$lib'"
for m in f16x3 bf16; do
DPN_LIB_OVERRIDE=$lib DPN_PHASE_DEBUG=1 timeout 120 python tools/step_jitter.py $m 12 2>&1 | grep -E "per-step|dpn phase" | tail -3 | cut -c1-260
done
done
