#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_balance.py -m gpu -x -q 2>&1 | tail -30
export MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_ctatimes.so
for b in 0 1; do timeout 200 python scripts/cta_times.py G16V 32 $b; done 2>&1 | tee gpurun_out/r3i_cta_times.txt
