#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export DPN_LIB_OVERRIDE=$PWD/tools/bin/libdpn_b200_debug.so
DPN_PHASE_DEBUG=1 timeout 90 python tools/step_jitter.py f16x3 3 2>&1 | grep -E "phase." | tail -2 | cut -c1-500 | tee gpurun_out/r02j_phase.txt
unset DPN_LIB_OVERRIDE
timeout 200 python tools/insitu_kernels.py 2>&1 | tail -25 | tee gpurun_out/r02j_insitu.txt
