cd /root/repo
timeout 300 python tools/chunk_ab.py f16x3 16384 32768 65536 2>&1 | tail -8 | tee gpurun_out/r02d_chunk_ab.txt
export DPN_LIB_OVERRIDE=$PWD/tools/bin/libdpn_b200_debug.so
DPN_PHASE_DEBUG=1 timeout 90 python tools/step_jitter.py f16x3 3 2>&1 | grep -E "phase." | tail -2 | cut -c1-500 | tee gpurun_out/r02d_phase.txt
