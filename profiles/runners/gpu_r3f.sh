#!/bin/bash
# how much of the high-D / small-batch sweep is table build: time against the number of frames
mkdir -p gpurun_out
for wl in "HIGHD 2" "HIGHD 4" "HIGHD 8" "G16V 2" "G16V 4" "G16V 16" "4CAM 2" "4CAM 8"; do
  timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1
done | tee gpurun_out/r3f_build_share.jsonl
