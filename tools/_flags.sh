#!/bin/bash
# debug build: pass 1 with work-skipping flags (1 = no MMAs, 4 = no tile stores) - in-situ kernel times + clocks
cd "$(dirname "$0")/.."
export DPN_LIB_OVERRIDE=$PWD/tools/bin/libdpn_b200_debug.so
for f in 0 1 4 5; do
  echo "== DPN_DEBUG_FLAGS=$f"
  DPN_DEBUG_FLAGS=$f timeout 200 python tools/insitu_kernels.py 2>&1 | grep -v -i Warn | head -4
done 2>&1 | tee gpurun_out/r02m_flags.txt
