#!/bin/bash
# high-D with the kept tap table: per-CTA timestamps and an ncu --set full capture of a steady-state USE launch
mkdir -p gpurun_out /tmp/rep
TAPTABLE=1 timeout 300 python scripts/cta_times.py HIGHD 2 0 2>&1 | tail -22 | tee gpurun_out/r4e_cta_times_highd.txt
BALANCE=0 TAPTABLE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 10 -c 1 -o /tmp/rep/prof_sweep_r4e python scripts/sweep_time.py HIGHD 2 8 > gpurun_out/ncu_sweep_r4e.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_sweep_r4e.ncu-rep gpurun_out/r4e_sweep_highd_ncu.txt --config "scripts/sweep_time.py HIGHD 2 (kept tap table, USE launch)" > /dev/null
cp /tmp/rep/prof_sweep_r4e.ncu-rep gpurun_out/
cat gpurun_out/r4e_sweep_highd_ncu.txt
