#!/bin/bash
# ncu --set full capture of the sweep kernel (GPU box).  Usage: bash profiles/runners/prof_sweep.sh <tag> [bench args]
tag=$1; shift
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o gpurun_out/prof_sweep_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_f.log 2>&1; tail -1 gpurun_out/ncu_f.log
