#!/bin/bash
# Evidence for profiles/: ncu launch list of one bench run, ncu --set full of the three tcgen05 kernels, accuracy of every mode.
#   TAG=r02 MODE=f16x3 gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-rXX}; MODE=${MODE:-f16x3}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --mode $MODE --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > gpurun_out/ncu_launch_stdout.txt 2>&1
for k in pass1_kernel pass2_kernel wgrad_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_$k \
      python bench.py --mode $MODE --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > gpurun_out/ncu_${k}_stdout.txt 2>&1
done
timeout 600 python tools/mode_accuracy.py > gpurun_out/mode_accuracy_$TAG.txt 2>&1
ls -la gpurun_out | tail -6
