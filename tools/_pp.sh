#!/bin/bash
# quick parity (split modes) + step time + in-situ kernel times of the in-tree library [and of variant builds: VARIANTS="a.so b.so"]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-quick}
timeout 400 python -m pytest tests/test_gpu_f16x3.py tests/test_gpu_bf16x3.py tests/test_gpu_decoder.py tests/test_gpu_parity.py tests/test_gpu_margin_fused.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_quick.txt
for lib in "" $VARIANTS; do
  echo "== library: ${lib:-in-tree}"
  DPN_LIB_OVERRIDE=${lib:+$PWD/$lib} timeout 200 python tools/insitu_kernels.py 2>&1 | grep -v -i Warn | head -4
done 2>&1 | tee gpurun_out/${T}_ab.txt
