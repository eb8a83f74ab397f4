#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 600 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "sweep_ms", round(r["kernel_ms"], 4), "frac", round(r["frac"], 3), "launches", d["gpu_launches"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2j_launches.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[h]
agg = {}
for r in rows[h + 1:]:
    d = dict(zip(cols, r))
    agg.setdefault(d["Kernel Name"][:60], []).append(float(d["Metric Value"].replace(",", "")))
for k, v in agg.items():
    print(f"{k:62s} n={len(v):3d} median {sorted(v)[len(v)//2]:.1f} {d['Metric Unit']}")
PY
