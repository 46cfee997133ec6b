#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -4
DPN_PHASE_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes 2>&1 >/dev/null | grep "dpn phase" | tail -2
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-modes 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f Mpts/s  %.3f ms/step  e2e %.2f Mpts/s %.2f ms  frac %.3f'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['e2e']['ms_per_step'],d['roofline']['frac']), d['clocks'])"
done
