#!/bin/bash
mkdir -p gpurun_out /tmp/rep
cat > /tmp/sa.py <<'PY'
import torch, os
from mvs_gi_b200 import modules, synthetic, ops
w = synthetic.WORKLOADS["G16V"]
dev = torch.device("cuda:0")
reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, w.scale, True).to(dev)
costs = [torch.randn((32, 1, 16, 80, 320), device=dev) for _ in range(4)]
for _ in range(3):
    for c in costs: reg.regress(c, want_norm_costs=False)
torch.cuda.synchronize()
PY
PYTHONPATH=$PWD ncu --set full --clock-control none --import-source on -k regex:softargmin -s 6 -c 1 -o /tmp/rep/prof_r2n_sa python /tmp/sa.py > gpurun_out/ncu_r2n.log 2>&1
cp /tmp/rep/prof_r2n_sa.ncu-rep gpurun_out/
ncu -i /tmp/rep/prof_r2n_sa.ncu-rep --page details 2>/dev/null | grep -E "Duration|Executed Ipc|Issue Slots Busy|Registers|Achieved Occupancy|Theoretical Occ|L1/TEX Hit|L2 Hit|DRAM Throughput|Memory Throughput|No Eligible|Stall|stall|Eligible Warps|Issued Warp|One or More|Local" | head -40
ncu -i /tmp/rep/prof_r2n_sa.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for a,b in zip(h,v):
    if 'issue_stalled' in a and 'per_issue_active' in a or 'pipe_xu' in a and 'pct' in a or 'pipe_alu.avg.pct_of_peak_sustained_active' in a or 'pipe_fma.avg.pct_of_peak_sustained_active' in a or 'pipe_lsu.avg.pct_of_peak_sustained_active' in a or 'inst_executed_pipe_xu' in a: print(a,b)
"
