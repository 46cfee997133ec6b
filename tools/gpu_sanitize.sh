#!/bin/bash
# compute-sanitizer on the tcgen05 kernels at small N: memcheck and racecheck, kernels of this library only.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --kernel-name kns=dpn --print-limit 20 --error-exitcode 0 \
      python tools/sanitize_case.py f16x3 bf16 > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "== $tool: exit $?"; grep -E "sanitize_case|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" gpurun_out/sanitizer_$tool.txt | head -12
done
