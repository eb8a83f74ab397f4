#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -4
for wl in "G16V 32" "G16V 8" "G16V 2" "4CAM 32" "G16VV_C48 16" "G16V_DS 32" "G8 32" "HIGHD 2"; do BALANCE=2 timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1; done | tee gpurun_out/r3s_build.jsonl
timeout 200 python scripts/cta_times.py G16V 32 1 2>&1 | grep "build_us"
