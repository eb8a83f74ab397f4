#!/bin/bash
# Round-2 evidence pass: parity tests, smoke, bench line (e2e + cpu baseline + reference on GPU), reference arm, launch
# list, ncu --set full of the sweep and the soft-argmin kernels.  Usage: bash profiles/runners/gpu_round2.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out /tmp/rep
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 1500 gpurun_out/bench_$tag.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 600 gpurun_out/bench_ref_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o /tmp/rep/prof_sweep_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ncu_sweep_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_sweep_$tag.ncu-rep gpurun_out/${tag}_sweep_ncu.txt --config "bench.py (defaults)" > /dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:softargmin -s 3 -c 1 -o /tmp/rep/prof_soft_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ncu_soft_$tag.log 2>&1
python scripts/ncu_summary.py /tmp/rep/prof_soft_$tag.ncu-rep gpurun_out/${tag}_softargmin_ncu.txt --config "bench.py (defaults)" > /dev/null
cp /tmp/rep/prof_sweep_$tag.ncu-rep /tmp/rep/prof_soft_$tag.ncu-rep gpurun_out/
head -12 gpurun_out/${tag}_sweep_ncu.txt
