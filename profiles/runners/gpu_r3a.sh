#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_softargmin_tma.py tests/test_gpu_parity.py tests/test_dropin.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python - <<'PY'
import torch, os
from mvs_gi_b200 import ops
dev = torch.device("cuda:0")
B, D, H, W = 32, 16, 80, 320
sets = [torch.randn((B, 1, D, H, W), device=dev) for _ in range(4)]
idx = torch.linspace(0.01, 1.0, D, device=dev)
def timeit(tensors, nc):
    for t in tensors: ops.softargmin_op(t, idx, 2 * H, 2 * W, 0.5, nc, False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(40): ops.softargmin_op(tensors[i % len(tensors)], idx, 2 * H, 2 * W, 0.5, nc, False)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 40
def mis(c):
    flat = torch.empty(c.numel() + 1, dtype=c.dtype, device=c.device); v = flat[1:].view(c.shape); v.copy_(c); return v
for nc in (False, True):
    print("norm_costs", nc, "tma ms", round(timeit(sets, nc), 4), "pointer ms", round(timeit([mis(s) for s in sets], nc), 4), flush=True)
PY
