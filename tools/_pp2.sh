#!/bin/bash
# cycle counters of pass 1 / pass 2 from the debug build (tools/build_debug.sh):  bash tools/_pp2.sh <tag>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-phase}
export DPN_LIB_OVERRIDE=$PWD/tools/bin/libdpn_b200_debug.so
DPN_PHASE_DEBUG=1 timeout 90 python tools/step_jitter.py f16x3 3 2>&1 | grep -E "phase." | tail -2 | cut -c1-500 | tee gpurun_out/${T}_phase.txt
