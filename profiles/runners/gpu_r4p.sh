#!/bin/bash
# final tree: GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_r4p.log 2>&1; tail -3 gpurun_out/pytest_gpu_r4p.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke_r4p.log
timeout 900 python bench.py > gpurun_out/bench_r4p.json 2> gpurun_out/bench_r4p.err; tail -c 300 gpurun_out/bench_r4p.json
