#!/bin/bash
# One GPU visit: parity tests of every mode, smoke, headline bench.   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json
tail -3 gpurun_out/bench_err.txt | cut -c1-300
