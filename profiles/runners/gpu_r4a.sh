#!/bin/bash
# Tap tables kept across calls: parity test, then the sweep kernel alone on every configuration with and without them.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tap_table.py -m gpu -x -q --timeout 300 2>&1 | tail -15
for wl in "G16V 32" "G16V 8" "G16V 2" "G16V 1" "4CAM 32" "G16VV_C48 16" "G16V_DS 32" "G8 32" "HIGHD 2" "HIGHD 8"; do
  for tt in 0 1; do BALANCE=2 TAPTABLE=$tt timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1; done
done | tee gpurun_out/r4a_tap_table.jsonl
