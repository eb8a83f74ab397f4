#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/sa_one.py <<'PY'
import torch
from mvs_gi_b200 import ops
dev = torch.device("cuda:0")
B, D, H, W = 2, 16, 80, 320
c = torch.randn((B, 1, D, H, W), device=dev)
idx = torch.linspace(0.01, 1.0, D, device=dev)
inv, _, _ = ops.softargmin_op(c, idx, 2 * H, 2 * W, 0.5, False, False)
torch.cuda.synchronize()
print("ok", float(inv.sum()))
PY
timeout 600 env PYTHONPATH=$PWD compute-sanitizer --tool memcheck python /tmp/sa_one.py > gpurun_out/sanitizer_sa.log 2>&1; head -30 gpurun_out/sanitizer_sa.log
