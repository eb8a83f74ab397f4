#!/bin/bash
# soft-argmin kernels: parity tests, then per-launch durations from ncu (the Python launch loop is slower than the kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_softargmin_tma.py -m gpu -x -q 2>&1 | tail -4
cat > /tmp/sa_dur.py <<'PY'
import torch
from mvs_gi_b200 import ops
dev = torch.device("cuda:0")
B, D, H, W = 32, 16, 80, 320
sets = [torch.randn((B, 1, D, H, W), device=dev) for _ in range(3)]
idx = torch.linspace(0.01, 1.0, D, device=dev)
def mis(c):
    flat = torch.empty(c.numel() + 1, dtype=c.dtype, device=c.device); v = flat[1:].view(c.shape); v.copy_(c); return v
msets = [mis(s) for s in sets]
for nc in (False, True):
    for group in (sets, msets):
        for i in range(9): ops.softargmin_op(group[i % 3], idx, 2 * H, 2 * W, 0.5, nc, False)
torch.cuda.synchronize()
PY
timeout 600 env PYTHONPATH=$PWD ncu --metrics gpu__time_duration.sum --clock-control none -k regex:softargmin --csv --log-file gpurun_out/r3d_sa.csv python /tmp/sa_dur.py > /dev/null 2>&1
python - <<'PY'
import csv, collections, statistics
rows = [r for r in csv.reader(open("gpurun_out/r3d_sa.csv")) if len(r) > 10]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]: d[r[ki][:70]].append(float(r[vi].replace(",", "")))
for k, v in d.items(): print(k, len(v), "median us", statistics.median(v) / 1000)
PY
