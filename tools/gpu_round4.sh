#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
