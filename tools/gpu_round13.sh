#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/step_jitter.py bf16 40 2>&1 | grep -E "per-step|CPU issue"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pass1|pass2|wgrad" -c 6 --csv --log-file gpurun_out/l2hint.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/l2hint.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[hi+1:]:
    if len(r)>vi: print(r[0], r[ki].split('(')[0][-16:], r[mi], r[vi], r[ui])
PY
