#!/bin/bash
# full re-measure: GPU tests, smoke, bench, ncu launch list, ncu --set full on the three tcgen05 kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r01b}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json
tail -3 gpurun_out/bench_err.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > gpurun_out/ncu_launch_stdout.txt 2>&1
for k in pass1_kernel pass2_kernel wgrad_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_$k \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > gpurun_out/ncu_${k}_stdout.txt 2>&1
done
ls -la gpurun_out | tail -8
