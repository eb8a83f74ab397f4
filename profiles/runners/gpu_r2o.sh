#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_full_size_parity.py -m gpu -x -q -k "softargmin or graph" 2>&1 | tail -3
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
for cfg in "base 2" "base 3" "base 4" "base 8"; do
  set -- $cfg
  PYTHONPATH=$PWD MVSGI_SOFTARGMIN_ROWS=$2 MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_$1.so ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:softargmin -c 24 --csv --log-file gpurun_out/r2o_sa.csv python /tmp/sa.py > /dev/null 2>&1
  python - "$cfg" <<'PY'
import csv, sys
rows = list(csv.reader(open("gpurun_out/r2o_sa.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[h]
agg = {}
for r in rows[h + 1:]:
    d = dict(zip(cols, r))
    agg.setdefault((d["Kernel Name"][-30:], d["Metric Name"][:24]), []).append(float(d["Metric Value"].replace(",", "")))
print(sys.argv[1], {k: round(sorted(v)[len(v)//2], 1) for k, v in agg.items()})
PY
done
