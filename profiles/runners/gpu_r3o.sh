#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_balance.py -m gpu -x -q 2>&1 | tail -12
for wl in "4CAM 32" "4CAM 8" "G16V 32" "HIGHD 2"; do for b in 0 1 2; do BALANCE=$b timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1; done; done | tee gpurun_out/r3o_balance_waves.jsonl
