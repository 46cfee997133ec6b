#!/bin/bash
# One GPU visit of round 2: the whole -m gpu suite, in-situ kernel times of the default mode, step time, headline bench.
#   gpurun --timeout 1500 -- 'bash tools/gpu_visit.sh [tag]'
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -80 > gpurun_out/${TAG}_pytest.txt; tail -12 gpurun_out/${TAG}_pytest.txt
for m in f16x3; do
  timeout 120 python tools/insitu_kernels.py $m 2>&1 | tail -7 | tee -a gpurun_out/${TAG}_insitu.txt
done
timeout 90 python tools/step_jitter.py f16x3 16 2>&1 | grep -E "per-step" | cut -c1-200 | tee -a gpurun_out/${TAG}_insitu.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 2> gpurun_out/${TAG}_bench_err.txt | tee gpurun_out/${TAG}_bench_n1.json | cut -c1-1500
tail -5 gpurun_out/${TAG}_bench_err.txt | cut -c1-300
