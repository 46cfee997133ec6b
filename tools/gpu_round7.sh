#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler.py tests/test_gpu_bf16.py -m gpu -x -q 2>&1 | tail -4
DPN_PHASE_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes 2>&1 >/dev/null | grep "dpn phase" | tail -4
