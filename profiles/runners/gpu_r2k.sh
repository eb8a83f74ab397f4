#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for rows in 2 3 4 6 9; do
MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_base.so MVSGI_SOFTARGMIN_ROWS=$rows python - <<PY
import torch, os
from mvs_gi_b200 import modules, synthetic, ops
w = synthetic.WORKLOADS["G16V"]
dev = torch.device("cuda:0")
reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, w.scale, True).to(dev)
costs = [torch.randn((32, 1, 16, 80, 320), device=dev) for _ in range(4)]
for nc in (False, True):
    for _ in range(5):
        for c in costs: reg.regress(c, want_norm_costs=nc)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        for c in costs: reg.regress(c, want_norm_costs=nc)
    b.record(); torch.cuda.synchronize()
    print("rows", os.environ["MVSGI_SOFTARGMIN_ROWS"], "norm_costs", nc, "us per 32 frames", round(a.elapsed_time(b) / 80 * 1000, 1))
PY
done
