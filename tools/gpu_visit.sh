#!/bin/bash
# One GPU visit of round 2: the whole -m gpu suite, in-situ kernel times of the default mode, step time.
#   gpurun --timeout 1500 -- 'bash tools/gpu_visit.sh [tag]'
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
for m in f16x3 bf16x3; do
  timeout 120 python tools/insitu_kernels.py $m 2>&1 | tail -7 | tee -a gpurun_out/${TAG}_insitu.txt
done
timeout 90 python tools/step_jitter.py f16x3 16 2>&1 | grep -E "per-step" | cut -c1-200 | tee -a gpurun_out/${TAG}_insitu.txt
