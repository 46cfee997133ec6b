#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_err.txt
