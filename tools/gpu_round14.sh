#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mb in 0 48 96; do
echo "== persist $mb MB"
DPN_L2_PERSIST_MB=$mb timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pass1|pass2|wgrad" -c 3 --csv --log-file gpurun_out/l2hint.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-modes 2>&1 >/dev/null | grep "dpn\]" | head -1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/l2hint.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[hi+1:]:
    if len(r)>vi: print(r[0], r[ki].split('(')[0][-16:], r[mi], r[vi], r[ui])
PY
DPN_L2_PERSIST_MB=$mb timeout 300 python tools/step_jitter.py bf16 30 2>&1 | grep -E "per-step"
done
