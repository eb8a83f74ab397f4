#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/sa.py <<'PY'
import torch, os
from mvs_gi_b200 import modules, synthetic, ops
w = synthetic.WORKLOADS["G16V"]
dev = torch.device("cuda:0")
reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, w.scale, True).to(dev)
costs = [torch.randn((32, 1, 16, 80, 320), device=dev) for _ in range(4)]
for nc in (False, True):
    for _ in range(3):
        for c in costs: reg.regress(c, want_norm_costs=nc)
torch.cuda.synchronize()
PY
for rows in 0 2 3 4 6; do
  if [ $rows = 0 ]; then export MVSGI_SOFTARGMIN_BLOCK=1; else unset MVSGI_SOFTARGMIN_BLOCK; export MVSGI_SOFTARGMIN_ROWS=$rows; fi
  PYTHONPATH=$PWD MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_base.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:softargmin --csv --log-file gpurun_out/r2l_sa_$rows.csv python /tmp/sa.py > /dev/null 2>&1
  python - $rows <<'PY'
import csv, sys
rows = list(csv.reader(open(f"gpurun_out/r2l_sa_{sys.argv[1]}.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[h]
agg = {}
for r in rows[h + 1:]:
    d = dict(zip(cols, r))
    agg.setdefault(d["Kernel Name"][:70], []).append(float(d["Metric Value"].replace(",", "")))
for k, v in agg.items():
    print("rows", sys.argv[1], f"{k:72s} n={len(v):3d} median {sorted(v)[len(v)//2]/1000:.1f} us")
PY
done
