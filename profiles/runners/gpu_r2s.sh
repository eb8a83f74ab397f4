#!/bin/bash
# FB (frames per pass) variants of sweep_kernel: time + checksum per library
mkdir -p gpurun_out
for v in base fb2_384x2 fb2_320x2 fb2_256x2 fb2_256x3 fb2_512x1 fb3_256x2 fb4_256x2 fb4_192x2; do
  for wl in "G16V 32" "4CAM 32"; do
    MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_$v.so timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1
  done
done | tee gpurun_out/r2s_fb_variants.jsonl
