#!/bin/bash
# One GPU iteration: parity tests, bench line, launch list.  Usage (on the GPU box): bash profiles/runners/gpu_cycle.sh <tag>
tag=${1:-x}
timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$tag.log 2>&1
tail -1 gpurun_out/bench_$tag.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fps',round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'sweep_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/launches_$tag.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:50]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in agg.items(): print(f"{k:52s} n={len(v):3d} mean_us={sum(v)/len(v)/1e3:9.1f}")
PY
