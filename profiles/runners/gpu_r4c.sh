#!/bin/bash
# ncu of the timed launches only (kept tap table, USE mode): launch list + --set full of one sweep launch, HIGHD timing
tag=r4c
mkdir -p gpurun_out /tmp/rep
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --profiler-range > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 1 -o /tmp/rep/prof_sweep_$tag python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --profiler-range > gpurun_out/ncu_sweep_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_sweep_$tag.ncu-rep gpurun_out/${tag}_sweep_ncu.txt --config "bench.py (defaults: kept tap table, USE launch inside the timed region)" > /dev/null
cp /tmp/rep/prof_sweep_$tag.ncu-rep gpurun_out/
head -40 gpurun_out/${tag}_sweep_ncu.txt
tail -5 gpurun_out/launches_$tag.csv
