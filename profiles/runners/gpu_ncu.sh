#!/bin/bash
# ncu --set full captures summarised ON the box (the .ncu-rep files are too large to bring back together); only the
# sweep report itself is kept.  Usage: bash profiles/runners/gpu_ncu.sh <tag>
tag=${1:-r1h}
mkdir -p gpurun_out /tmp/rep
CFG="bench.py default: G16V std, B=32 frames, per-frame float masks, NCHW feats, in-kernel projection, channels-last volume"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o /tmp/rep/prof_sweep_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ncu_sweep_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_sweep_$tag.ncu-rep gpurun_out/${tag}_sweep_ncu.txt --json gpurun_out/sweep_traffic.json --config "$CFG" > /dev/null
cp /tmp/rep/prof_sweep_$tag.ncu-rep gpurun_out/
timeout 900 ncu --set full --clock-control none -k regex:resample_kernel -s 3 -c 1 -o /tmp/rep/prof_resample_$tag python scripts/resample_bench.py 24 > gpurun_out/ncu_resample_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_resample_$tag.ncu-rep gpurun_out/${tag}_resample_ncu.txt --config "scripts/resample_bench.py 24: 24 images 1216x1216x3 f32 -> 512x2048" > /dev/null
timeout 900 ncu --set full --clock-control none -k regex:sweep_backward_kernel -s 3 -c 1 -o /tmp/rep/prof_bwd_$tag python scripts/backward_bench.py 16 > gpurun_out/ncu_bwd_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_bwd_$tag.ncu-rep gpurun_out/${tag}_sweep_backward_ncu.txt --config "scripts/backward_bench.py 16: G16V std, 16 frames" > /dev/null
timeout 900 ncu --set full --clock-control none -k regex:softargmin_backward -s 3 -c 1 -o /tmp/rep/prof_sbwd_$tag python scripts/backward_bench.py 16 > gpurun_out/ncu_sbwd_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_sbwd_$tag.ncu-rep gpurun_out/${tag}_softargmin_backward_ncu.txt --config "scripts/backward_bench.py 16: G16V, 16 frames" > /dev/null
timeout 600 python bench.py > gpurun_out/bench_$tag.log 2>&1; tail -1 gpurun_out/bench_$tag.log | cut -c1-900
ls -la gpurun_out | tail -12; head -30 gpurun_out/${tag}_sweep_ncu.txt
