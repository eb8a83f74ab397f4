#!/bin/bash
# Full GPU pass (on the GPU box): parity tests, smoke, bench line (with e2e + cpu baseline), launch list, ncu --set full of
# the sweep kernel and the soft-argmin kernel.  Usage: bash profiles/runners/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_$tag.csv &
SMI=$!
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.log 2>&1; tail -1 gpurun_out/bench_$tag.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.log 2>&1; tail -1 gpurun_out/bench_ref_$tag.log
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o gpurun_out/prof_sweep_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_sweep_$tag.log 2>&1; tail -1 gpurun_out/ncu_sweep_$tag.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:softargmin -s 3 -c 1 -o gpurun_out/prof_soft_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_soft_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_feats -s 3 -c 1 -o gpurun_out/prof_stage_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_stage_$tag.log 2>&1
ls -la gpurun_out
