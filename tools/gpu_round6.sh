#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.txt
timeout 600 python tools/e2e_profile.py > gpurun_out/e2e_profile.txt 2>&1; tail -45 gpurun_out/e2e_profile.txt | cut -c1-220
for k in pass1_kernel pass2_kernel wgrad_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof2_$k \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > gpurun_out/ncu_${k}_stdout.txt 2>&1
done
ls -la gpurun_out | tail -5
