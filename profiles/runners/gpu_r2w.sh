#!/bin/bash
# ncu --set full of the shipped sweep kernel at the bench shape (G16V, 32 frames, static masks, in-place features)
mkdir -p gpurun_out /tmp/rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 1 -o /tmp/rep/prof_r2w_sweep python scripts/sweep_time.py G16V 32 4 > gpurun_out/ncu_r2w.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_r2w_sweep.ncu-rep gpurun_out/r2w_sweep_ncu.txt --config "scripts/sweep_time.py G16V 32" > /dev/null
cp /tmp/rep/prof_r2w_sweep.ncu-rep gpurun_out/
cat gpurun_out/r2w_sweep_ncu.txt | head -50
