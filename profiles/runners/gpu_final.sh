#!/bin/bash
# Round-end evidence pass: tests, smoke, both bench arms, launch list, ncu --set full summary of the sweep kernel.
tag=${1:-r1k}
mkdir -p gpurun_out /tmp/rep
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.log 2>&1; tail -1 gpurun_out/bench_$tag.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.log 2>&1; tail -1 gpurun_out/bench_ref_$tag.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > /dev/null 2>&1
CFG="bench.py default: G16V std, B=32 frames, per-frame float masks, NCHW feats, in-kernel projection, channels-last volume"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o /tmp/rep/prof_sweep_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ncu_sweep_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_sweep_$tag.ncu-rep gpurun_out/${tag}_sweep_ncu.txt --json gpurun_out/sweep_traffic.json --config "$CFG" | head -12
cp /tmp/rep/prof_sweep_$tag.ncu-rep gpurun_out/
