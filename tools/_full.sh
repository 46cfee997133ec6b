#!/bin/bash
# full GPU suite + smoke + in-situ kernel times (no bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02j}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/${T}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${T}_smoke.txt
