#!/bin/bash
# full GPU suite + smoke + headline bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-rXX}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${T}_smoke.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 2> gpurun_out/${T}_bench_err.txt | tee gpurun_out/${T}_bench_n1.json | cut -c1-1500
tail -3 gpurun_out/${T}_bench_err.txt | cut -c1-300
