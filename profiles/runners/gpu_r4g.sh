#!/bin/bash
# L2 prefetch of the next item's table + wide frames-per-pass variant from 4 frames on when the table is kept
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tap_table.py tests/test_balance.py -m gpu -x -q --timeout 300 2>&1 | tail -5
{
BALANCE=0 TAPTABLE=1 timeout 200 python scripts/sweep_time.py HIGHD 2 2>&1 | tail -1
BALANCE=0 TAPTABLE=1 timeout 200 python scripts/sweep_time.py HIGHD 8 2>&1 | tail -1
BALANCE=0 TAPTABLE=1 timeout 200 python scripts/sweep_time.py HIGHD 4 2>&1 | tail -1
BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py 4CAM 32 2>&1 | tail -1
BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py 4CAM 4 2>&1 | tail -1
BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py G16V 32 2>&1 | tail -1
BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py G16V 8 2>&1 | tail -1
BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py G16V 4 2>&1 | tail -1
BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py G8 8 2>&1 | tail -1
} | tee gpurun_out/r4g_prefetch_fb.jsonl
