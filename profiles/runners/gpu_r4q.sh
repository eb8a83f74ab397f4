#!/bin/bash
# ncu --set full of steady-state USE launches: high-D (item queue) and C = 48
mkdir -p gpurun_out /tmp/rep
BALANCE=0 TAPTABLE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 10 -c 1 -o /tmp/rep/prof_highd python scripts/sweep_time.py HIGHD 2 8 > gpurun_out/ncu_r4q_a.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_highd.ncu-rep gpurun_out/r4q_sweep_highd_ncu.txt --config "scripts/sweep_time.py HIGHD 2 (kept tap table, item queue, USE launch)" > /dev/null
BALANCE=1 TAPTABLE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 12 -c 1 -o /tmp/rep/prof_c48 python scripts/sweep_time.py G16VV_C48 16 8 > gpurun_out/ncu_r4q_b.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_c48.ncu-rep gpurun_out/r4q_sweep_c48_ncu.txt --config "scripts/sweep_time.py G16VV_C48 16 (kept tap table, balanced runs, USE launch)" > /dev/null
grep -h "kernel:\|gpu__time_duration\|dram__bytes\|lts__t_sector_hit\|l1tex__data_pipe_lsu_wavefronts.avg\|l1tex__t_sector_hit\|long_scoreboard\|stalled_barrier\|issue_active" gpurun_out/r4q_sweep_highd_ncu.txt gpurun_out/r4q_sweep_c48_ncu.txt
