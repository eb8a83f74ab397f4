#!/bin/bash
# Item queue (kept table, more items than CTAs) and rotated candidate chunks: parity tests, then high-D / 4 cameras timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tap_table.py tests/test_full_size_parity.py -m gpu -x -q --timeout 300 2>&1 | tail -5
{
for tt in 0 1; do BALANCE=0 TAPTABLE=$tt timeout 200 python scripts/sweep_time.py HIGHD 2 2>&1 | tail -1; done
for tt in 0 1; do BALANCE=0 TAPTABLE=$tt timeout 200 python scripts/sweep_time.py HIGHD 8 2>&1 | tail -1; done
BALANCE=0 TAPTABLE=1 timeout 200 python scripts/sweep_time.py HIGHD 1 2>&1 | tail -1
BALANCE=0 TAPTABLE=1 timeout 200 python scripts/sweep_time.py 4CAM 32 2>&1 | tail -1
BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py 4CAM 32 2>&1 | tail -1
BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py G16V 32 2>&1 | tail -1
} | tee gpurun_out/r4f_item_queue.jsonl
TAPTABLE=1 timeout 300 python scripts/cta_times.py HIGHD 2 0 2>&1 | grep '"workload"' | tee gpurun_out/r4f_cta_times_highd.txt
