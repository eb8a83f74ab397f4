#!/bin/bash
# high-D with the kept table: candidates per chunk 8 / 16 (default) / 32 / 64 (variant libraries built with another default)
mkdir -p gpurun_out
for d in 8 16 32 64; do
  lib=$PWD/mvs_gi_b200/libmvsgi_b200_dch$d.so; [ $d = 16 ] && lib=$PWD/mvs_gi_b200/libmvsgi_b200.so
  for wl in "HIGHD 2" "HIGHD 8"; do MVSGI_B200_LIB=$lib BALANCE=0 TAPTABLE=1 timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1 | sed "s/^{/{\"dch\": $d, /"; done
done | tee gpurun_out/r4m_highd_dch.jsonl
