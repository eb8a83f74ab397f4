#!/bin/bash
# Final round evidence: full GPU suite, smoke, bench line, reference arm; ncu of the timed launches only.
tag=r4h
mkdir -p gpurun_out /tmp/rep
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 300 gpurun_out/bench_ref_$tag.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --profiler-range > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 1 -o /tmp/rep/prof_sweep_$tag python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --profiler-range > gpurun_out/ncu_sweep_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_sweep_$tag.ncu-rep gpurun_out/${tag}_sweep_ncu.txt --config "bench.py (defaults: kept tap table, USE launch inside the timed region)" > /dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:softargmin -s 2 -c 1 -o /tmp/rep/prof_soft_$tag python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --profiler-range > gpurun_out/ncu_soft_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_soft_$tag.ncu-rep gpurun_out/${tag}_softargmin_ncu.txt --config "bench.py (defaults)" > /dev/null
cp /tmp/rep/prof_sweep_$tag.ncu-rep gpurun_out/
for wl in "G16V 32" "G16V 8" "G16V 2" "G16V 1" "4CAM 32" "G16VV_C48 16" "G16V_DS 32" "G8 32" "HIGHD 2" "HIGHD 8"; do BALANCE=2 TAPTABLE=1 timeout 200 python scripts/sweep_time.py $wl 2>&1 | tail -1; done | tee gpurun_out/${tag}_all_configs.jsonl
