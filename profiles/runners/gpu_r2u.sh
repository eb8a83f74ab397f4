#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_inplace.py tests/test_gpu_parity.py tests/test_full_size_parity.py tests/test_dropin.py -m gpu -x -q 2>&1 | tail -5
for wl in "G16V 32" "G16V 8" "G16V 2" "4CAM 32" "G16VV_C48 16" "HIGHD 2" "G16V_DS 32" "G8 32" "E8 32"; do
  timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1
done | tee gpurun_out/r2u_default_lib.jsonl
