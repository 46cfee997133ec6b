#!/bin/bash
# Round 2, first visit: parity of the N-half pipeline + double-buffered wgrad, then in-situ kernel times of both pass-1 variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_f16x3.py tests/test_gpu_bf16x3.py tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.txt
timeout 200 python -m pytest tests/test_gpu_pass1_variants.py tests/test_gpu_bf16.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 | tee -a gpurun_out/r2a_pytest.txt
for v in np ts; do
  DPN_P1=$v timeout 120 python tools/insitu_kernels.py f16x3 2>&1 | tail -8 | tee -a gpurun_out/r2a_insitu.txt
done
timeout 120 python tools/insitu_kernels.py bf16 2>&1 | tail -8 | tee -a gpurun_out/r2a_insitu.txt
timeout 90 python tools/step_jitter.py f16x3 16 2>&1 | grep -E "per-step" | cut -c1-200 | tee -a gpurun_out/r2a_insitu.txt
