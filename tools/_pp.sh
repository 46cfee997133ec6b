#!/bin/bash
# ping-pong pipelined pass 1 / pass 2 against the previous commit's library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_f16x3.py tests/test_gpu_bf16x3.py tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02j_pytest_quick.txt
for lib in "" "$PWD/tools/bin/libdpn_prev.so"; do
  echo "== library: ${lib:-in-tree}"
  DPN_LIB_OVERRIDE=$lib timeout 120 python tools/step_jitter.py f16x3 16 2>&1 | grep -E "per-step" | cut -c1-200
done 2>&1 | tee gpurun_out/r02j_ab.txt
