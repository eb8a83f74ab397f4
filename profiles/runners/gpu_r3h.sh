#!/bin/bash
# cost-balanced runs: parity, then sweep time with and without the table on every configuration
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_balance.py tests/test_inplace.py tests/test_dropin.py -m gpu -x -q 2>&1 | tail -5
for wl in "G16V 32" "G16V 8" "4CAM 32" "G16VV_C48 16" "G16V_DS 32" "G8 32" "HIGHD 2"; do
  for bal in 0 1; do
    BALANCE=$bal timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1
  done
done | tee gpurun_out/r3j_balance.jsonl
