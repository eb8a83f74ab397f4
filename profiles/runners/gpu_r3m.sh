#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_balance.py -m gpu -x -q 2>&1 | tail -15
for m in 2; do timeout 200 python scripts/cta_times.py G16V 32 $m 2>&1 | grep -v "^  SM"; done | tee gpurun_out/r3m_cta_times.txt
for wl in "G16V 32" "G16V 8" "G16VV_C48 16" "G16V_DS 32"; do BALANCE=2 timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1; done | tee gpurun_out/r3m_calibrated.jsonl
