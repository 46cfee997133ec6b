#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/tc_debug.py 300 1 > gpurun_out/tc_debug.txt 2>&1; echo "tc_debug rc=$?"; head -c 4500 gpurun_out/tc_debug.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json
tail -3 gpurun_out/bench_err.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-modes > gpurun_out/ncu_launch_stdout.txt 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; data=rows[hi+1:]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in data:
    if len(r)<=vi: continue
    name=r[ki].split('(')[0]
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e3 if u=='us' else v/1e6 if u=='ns' else v*1e3 if u=='s' else v
    agg.setdefault(name,[]).append(v)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:8]:
    print("%-50s %4d %9.3f ms total %8.3f mean"%(k[:50],len(v),sum(v),sum(v)/len(v)))
PY
