#!/bin/bash
# Round 2, second visit: new rows (fused margin, query generator, sampler bounds, headline parity) + where pass1_np spends its time.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_margin_fused.py tests/test_query_generator.py tests/test_sampler.py tests/test_gpu_trainer.py tests/test_gpu_graphed.py tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2b_pytest.txt
timeout 600 python -m pytest tests/test_gpu_headline_parity.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -70 | tee gpurun_out/r2b_headline.txt
export DPN_LIB_OVERRIDE=$PWD/tools/bin/libdpn_b200_debug.so
DPN_PHASE_DEBUG=1 timeout 90 python tools/step_jitter.py f16x3 4 2>&1 | grep -E "phase. pass1_np|phase. pass2|per-step" | tail -3 | cut -c1-400 | tee gpurun_out/r2b_phase.txt
for fl in 0 1 2 4 6; do
  echo "== DPN_DEBUG_FLAGS=$fl" | tee -a gpurun_out/r2b_phase.txt
  DPN_DEBUG_FLAGS=$fl timeout 120 python tools/insitu_kernels.py f16x3 2>&1 | grep -E "pass1|pass2|wgrad" | tee -a gpurun_out/r2b_phase.txt
done
DPN_P1=ts DPN_PHASE_DEBUG=1 timeout 90 python tools/step_jitter.py f16x3 4 2>&1 | grep -E "phase. pass1|per-step" | tail -3 | cut -c1-400 | tee -a gpurun_out/r2b_phase.txt
