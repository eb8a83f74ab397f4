#!/bin/bash
mkdir -p gpurun_out
for v in "" _pf _pf_ef; do
  for wl in "G16V 32" "G16V 8" "4CAM 32" "G16VV_C48 16" "HIGHD 2" "G16V_DS 32"; do
    MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200$v.so timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1
  done
done | tee gpurun_out/r2x_prefetch.jsonl
