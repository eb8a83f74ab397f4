#!/bin/bash
# soft-argmin backward as a gather: parity, then per-launch durations from ncu and the backward bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "softargmin" 2>&1 | tail -4
cat > /tmp/sb_dur.py <<'PY'
import torch
from mvs_gi_b200 import ops
dev = torch.device("cuda:0")
B, D, H, W = 16, 16, 80, 320
sets = [torch.randn((B, 1, D, H, W), device=dev) for _ in range(3)]
idx = torch.linspace(0.01, 1.0, D, device=dev)
gi = torch.randn((B, 1, 2 * H, 2 * W), device=dev)
gn = [torch.randn((B, D, 2 * H, 2 * W), device=dev) for _ in range(3)]
for i in range(9): ops.softargmin_backward_op(gi, None, sets[i % 3], idx, 2 * H, 2 * W, 0.5)
for i in range(9): ops.softargmin_backward_op(gi, gn[i % 3], sets[i % 3], idx, 2 * H, 2 * W, 0.5)
torch.cuda.synchronize()
PY
timeout 600 env PYTHONPATH=$PWD ncu --metrics gpu__time_duration.sum --clock-control none -k regex:softargmin --csv --log-file gpurun_out/r3e_sb.csv python /tmp/sb_dur.py > /dev/null 2>&1
python - <<'PY'
import csv, collections, statistics
rows = [r for r in csv.reader(open("gpurun_out/r3e_sb.csv")) if len(r) > 10]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]: d[r[ki][:70]].append(float(r[vi].replace(",", "")))
for k, v in d.items(): print(k, len(v), "median us", statistics.median(v) / 1000)
PY
timeout 300 python scripts/backward_bench.py 16 | tee gpurun_out/r3e_backward_bench.json
