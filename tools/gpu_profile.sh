#!/bin/bash
# Evidence for profiles/: ncu launch list of one bench run, ncu --set full of the three tcgen05 kernels, sanitizer logs.
#   TAG=r02 MODE=f16x3 gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-rXX}; MODE=${MODE:-f16x3}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --mode $MODE --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes --no-configs --no-eager-gpu --no-sustained > gpurun_out/ncu_launch_stdout.txt 2>&1
for k in pass1_ts_kernel pass2z_kernel wgrad2_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_$k \
      python bench.py --mode $MODE --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes --no-configs --no-eager-gpu --no-sustained > gpurun_out/ncu_${k}_stdout.txt 2>&1
done
bash tools/gpu_sanitize.sh
ls -la gpurun_out | tail -8
