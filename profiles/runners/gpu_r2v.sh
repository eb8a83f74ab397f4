#!/bin/bash
mkdir -p gpurun_out
{
for v in abl_l1 abl_nost abl_both; do
  MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_$v.so timeout 200 python scripts/sweep_time.py G16V 32 2>&1 | tail -1
done
SAMEFRAME=1 timeout 200 python scripts/sweep_time.py G16V 32 2>&1 | tail -1
SAMEFRAME=1 MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_abl_nost.so timeout 200 python scripts/sweep_time.py G16V 32 2>&1 | tail -1
} | tee gpurun_out/r2v_ablation.jsonl
