#!/bin/bash
# ncu --set full of the sweep kernel(s), summarised on the box.  Usage: bash profiles/runners/gpu_ncu2.sh <tag> <kernel regex> [bench args]
tag=$1; regex=$2; shift 2
mkdir -p gpurun_out /tmp/rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s 3 -c 1 -o /tmp/rep/prof_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants "$@" > gpurun_out/ncu_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_$tag.ncu-rep gpurun_out/${tag}_ncu.txt --config "bench.py $*" > /dev/null
cp /tmp/rep/prof_$tag.ncu-rep gpurun_out/
cat gpurun_out/${tag}_ncu.txt
