#!/bin/bash
mkdir -p gpurun_out
for v in base fb2_256x2 fb3_256x2 fb4_256x2 fb2_384x2; do
  for wl in "G16V 32" "G16V 8" "G16V 2" "4CAM 32" "G16VV_C48 16" "HIGHD 2" "HIGHD 8" "G16V_DS 32"; do
    MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_$v.so timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1
  done
done | tee gpurun_out/r2t_fb_variants.jsonl
