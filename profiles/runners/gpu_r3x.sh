#!/bin/bash
mkdir -p gpurun_out
for cfg in "3 3 0.015" "6 5 0.015" "6 5 0.008" "10 5 0.01"; do set -- $cfg
  for rep in 1 2; do CAL_ROUNDS=$1 CAL_LAUNCHES=$2 CAL_THRESHOLD=$3 BALANCE=2 timeout 200 python scripts/sweep_time.py G16V 32 60 2>&1 | tail -2 | tr '\n' ' '; echo " <- rounds $1 launches $2 thr $3"; done
done | tee gpurun_out/r3x_calibration_params.txt
