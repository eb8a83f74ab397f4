#!/bin/bash
# kept table, fewer than 16 frames: frames per pass picked by cost
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tap_table.py tests/test_balance.py tests/test_inplace.py -m gpu -x -q --timeout 300 2>&1 | tail -3
for wl in "G16V 2" "G16V 3" "G16V 4" "G16V 5" "G16V 6" "G16V 7" "G16V 8" "G16V 12" "4CAM 3" "4CAM 4" "4CAM 5" "G16VV_C48 3" "HIGHD 2"; do BALANCE=1 TAPTABLE=1 timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1; done | tee gpurun_out/r4o_fb_by_cost.jsonl
