#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_bf16x3.py tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pass1|pass2|wgrad" -c 3 --csv --log-file gpurun_out/x3.csv \
    python bench.py --mode bf16x3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/x3.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[hi+1:]:
    if len(r)>vi: print(r[0], r[ki].split('(')[0][-24:], r[mi], r[vi], r[ui])
PY
DPN_PHASE_DEBUG=1 timeout 120 python bench.py --mode bf16x3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes 2>&1 >/dev/null | grep "dpn phase" | tail -2
