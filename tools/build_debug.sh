#!/bin/bash
# Debug build of the library for timing experiments: tools/bin/libdpn_b200_debug.so (-DDPN_DEBUG_BUILD: cycle counters with
# DPN_PHASE_DEBUG=1, work-skipping switches with DPN_DEBUG_FLAGS).  Use with DPN_LIB_OVERRIDE=tools/bin/libdpn_b200_debug.so.
# The shipped library (__graft_entry__.build) never contains these.
cd "$(dirname "$0")/.."
mkdir -p tools/bin
S=deepphysinet_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DDPN_DEBUG_BUILD=1 "$@" \
    $S/dpn_api.cu $S/dpn_fp32.cu $S/dpn_tc.cu $S/dpn_sampler.cu -o tools/bin/libdpn_b200_debug.so
