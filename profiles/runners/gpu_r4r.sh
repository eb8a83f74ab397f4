#!/bin/bash
# groups of 8 records handed to the warps through a shared-memory counter: full GPU suite, then every configuration
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_r4r.log 2>&1; tail -3 gpurun_out/pytest_gpu_r4r.log
for wl in "G16V 32" "G16V 8" "G16V 1" "4CAM 32" "G16VV_C48 16" "G16V_DS 32" "G8 32" "HIGHD 2"; do BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1; done | tee gpurun_out/r4r_group_counter.jsonl
BALANCE=2 TAPTABLE=0 timeout 200 python scripts/sweep_time.py "G16V" 32 2>&1 | tail -1 | tee -a gpurun_out/r4r_group_counter.jsonl
