#!/bin/bash
# kept table, fewer than 4 frames: one frame per pass x 512 threads (variant build) against two frames per pass x 384
mkdir -p gpurun_out
for v in default fb1; do
  lib=$PWD/mvs_gi_b200/libmvsgi_b200_fb1.so; [ $v = default ] && lib=$PWD/mvs_gi_b200/libmvsgi_b200.so
  for wl in "HIGHD 2" "G16V 2" "G16V 3" "4CAM 2"; do MVSGI_B200_LIB=$lib BALANCE=1 TAPTABLE=1 timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1; done
done | tee gpurun_out/r4n_small_batch_fb1.jsonl
