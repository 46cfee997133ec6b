#!/bin/bash
# cross-first accumulation (in-tree) against the interleaved order (tools/bin/libdpn_noxfirst.so): parity, seed sweep, kernel times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-xf}
timeout 400 python -m pytest tests/test_gpu_f16x3.py tests/test_gpu_bf16x3.py tests/test_gpu_decoder.py tests/test_gpu_parity.py tests/test_gpu_margin_fused.py -m gpu -x -q 2>&1 | tail -6
for lib in ${LIBS:-"" tools/bin/libdpn_noxfirst.so}; do
  echo "== library: ${lib:-in-tree (cross-first)}"
  DPN_LIB_OVERRIDE=${lib:+$PWD/$lib} timeout 300 python -m pytest tests/test_gpu_headline_parity.py -m gpu -x -q -s 2>&1 | grep -E "f16x3|draws|headline" | cut -c1-250
  DPN_LIB_OVERRIDE=${lib:+$PWD/$lib} timeout 200 python tools/insitu_kernels.py 2>&1 | grep -v -i Warn | head -4
done 2>&1 | tee gpurun_out/${T}_ab.txt
