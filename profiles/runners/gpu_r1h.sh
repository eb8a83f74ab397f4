#!/bin/bash
# Final GPU pass of the round (on the GPU box): tests, smoke, bench lines (both arms), launch list, ncu --set full of
# the sweep / resample / backward kernels, backward + resample measurements.  Usage: bash profiles/runners/gpu_r1h.sh <tag>
tag=${1:-r1h}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_$tag.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.log 2>&1; tail -1 gpurun_out/bench_$tag.log
timeout 600 python bench.py --vol-layout ncdhw --no-cpu-baseline > gpurun_out/bench_${tag}_ncdhw.log 2>&1; tail -1 gpurun_out/bench_${tag}_ncdhw.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.log 2>&1; tail -1 gpurun_out/bench_ref_$tag.log
timeout 300 python scripts/backward_bench.py 16 > gpurun_out/backward_bench_$tag.json 2> gpurun_out/backward_bench_$tag.err; cat gpurun_out/backward_bench_$tag.json; tail -2 gpurun_out/backward_bench_$tag.err
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o gpurun_out/prof_sweep_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ncu_sweep_$tag.log 2>&1; tail -1 gpurun_out/ncu_sweep_$tag.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resample_kernel -s 3 -c 1 -o gpurun_out/prof_resample_$tag python scripts/resample_bench.py 24 > gpurun_out/ncu_resample_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_backward_kernel -s 3 -c 1 -o gpurun_out/prof_bwd_$tag python scripts/backward_bench.py 16 > gpurun_out/ncu_bwd_$tag.log 2>&1
ls -la gpurun_out | tail -20
