#!/bin/bash
mkdir -p gpurun_out
for wl in "G16V 32" "HIGHD 2" "4CAM 32"; do BALANCE=1 timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1; done
timeout 200 python scripts/cta_times.py G16V 32 1 2>&1 | grep "build_us"
