#!/bin/bash
mkdir -p gpurun_out
export MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_ctatimes.so
timeout 200 python scripts/cta_times.py G16V 32 1 2>&1 | grep -v "^  SM" | tee gpurun_out/r3l_cta_times.txt
